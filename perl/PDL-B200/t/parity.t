#!/usr/bin/env perl
# Same-process parity: every op runs twice on the same ndarrays — through libpdlb200
# (PDL::B200 attached) and through the reference's own CPU readdata (PDL::B200::enable(0)) —
# and the two results must agree in type, dims, badflag and BYTES.
use strict; use warnings;
use Test::More;
use PDL::LiteF;
use PDL::B200;

plan skip_all => 'no CUDA device' unless PDL::B200::device_count() > 0;

my $seed = 4242;
sub rnd { $seed = ($seed * 1103515245 + 12345) % 2147483648; $seed }
sub mk {
  my ($type, @dims) = @_;
  my $n = 1; $n *= $_ for @dims;
  my $p = pdl($type, [map { (rnd() % 2001 - 1000) / ($type->integer ? 1 : 8) } 1 .. $n]);
  $p = abs($p) if $type->unsigned;
  return @dims > 1 ? $p->reshape(@dims) : $p;
}
sub bytes { my $q = $_[0]->copy; $q->make_physical; unpack('H*', ${ $q->get_dataref }) }
sub both {
  my ($name, $code) = @_;
  PDL::B200::enable(1); my $g = $code->();
  PDL::B200::enable(0); my $c = $code->();
  PDL::B200::enable(1);
  is($g->type . '', $c->type . '', "$name: type");
  is_deeply([$g->dims], [$c->dims], "$name: dims");
  is($g->badflag, $c->badflag, "$name: badflag");
  is(bytes($g), bytes($c), "$name: bytes");
}

my @types = (byte, short, ushort, long, indx, longlong, float, double);
for my $t (@types) {
  my ($a, $b) = (mk($t, 300, 7), mk($t, 300, 7));
  my $bp = $b->copy; $bp->where($bp == 0) .= 3;
  both("plus $t",   sub { $a + $b });
  both("minus $t",  sub { $a - $b });
  both("mult $t",   sub { $a * $b });
  both("divide $t", sub { $a / $bp });
  both("gt $t",     sub { $a > $b });
  both("modulo $t", sub { $a % $bp });
  both("sumover $t",     sub { $a->sumover });
  both("average $t",     sub { $a->average });
  both("minimum $t",     sub { $a->minimum });
  both("maximum_ind $t", sub { $a->maximum_ind });
  both("sum of xchg $t", sub { $a->xchg(0, 1)->sumover });
  both("slice+dummy $t", sub { $a->slice('0:-1:3,(2)')->dummy(1, 4) + $b->slice('5:104,0:3') });
  both("scalar promote $t", sub { $a + 1.5 });
  if (!$t->integer || $t == long) {
    my ($ma, $mb) = (mk($t, 40, 17), mk($t, 9, 40));     # exact kernel: same summation order => same bytes
    both("matmult $t", sub { $ma x $mb });
  }
  my $ip = $a->copy;
  both("inplace plus $t", sub { my $q = $ip->copy; $q->inplace->plus($b, 0); $q });
  both("+= small $t", sub { my $q = mk($t, 5)->copy * 0 + 3; $q += 4; $q });
  my $bad = $a->copy; $bad->badflag(1); $bad->flat->setbadat($_) for (0, 17, 299, 300 .. 599);
  both("bad plus $t",    sub { $bad + $b });
  both("bad sumover $t", sub { $bad->sumover });
  both("bad average $t", sub { $bad->average });
  both("bad minimum $t", sub { $bad->minimum });
  # the rows either side of the path (SURVEY.md §8(f)): Bad.pd ops, constructors, scans, inner
  my $mask = ($a->abs % 5 == 0);
  both("isbad $t",        sub { $bad->isbad });
  both("isgood $t",       sub { $bad->isgood });
  both("setbadif $t",     sub { $a->setbadif($mask) });
  both("setbadif+sum $t", sub { $a->setbadif($mask)->sumover });
  both("setvaltobad $t",  sub { $a->setvaltobad(7) });
  both("setbadtoval $t",  sub { $bad->setbadtoval(3) });
  both("setbadtoval inplace $t", sub { my $q = $bad->copy; $q->inplace->setbadtoval(1); $q });
  both("copybad $t",      sub { $b->copybad($bad) });
  both("badmask $t",      sub { $bad->badmask(5) });
  both("nbadover $t",     sub { $bad->nbadover });
  both("cumusumover $t",  sub { ($a % 7)->cumusumover });
  both("bad cumusumover $t", sub { $bad->cumusumover });
  both("xvals $t",        sub { $a->xvals });
  both("yvals $t",        sub { $a->yvals });
  both("sequence $t",     sub { sequence($t, 301, 5) });
  both("inner $t",        sub { inner($a % 3, $b % 3) });
  both("inner bad $t",    sub { inner($bad % 3, $b % 3) });
  for my $k (0 .. 3) {
    both("minmaximum[$k] $t",     sub { ($a->minmaximum)[$k] });
    both("bad minmaximum[$k] $t", sub { ($bad->minmaximum)[$k] });
  }
  is_deeply([do { PDL::B200::enable(1); $a->minmax }], [do { PDL::B200::enable(0); my @r = $a->minmax; PDL::B200::enable(1); @r }], "minmax $t");
  both("magnover $t",     sub { ($a % 5)->magnover });
  both("outer $t",        sub { outer($a->slice(':,(1)') % 9, $b->slice('0:40,(2)') % 9) });
  both("outer bad $t",    sub { outer($bad->slice(':,(0)') % 9, $b->slice('0:40,(2)') % 9) });
  both("minimum_n_ind $t",     sub { $a->minimum_n_ind(4) });
  both("maximum_n_ind bad $t", sub { $bad->maximum_n_ind(3) });
  both("maximum_n_ind short $t", sub { $bad->slice('0:1,:')->maximum_n_ind(2) });     # all-BAD rows: BAD slots + badflag
  both("abs2 $t",         sub { ($a % 11)->abs2 });
  both("convert $t -> float", sub { $a->float });
  both("convert bad $t -> long", sub { $bad->long });
  both("flat sum $t",     sub { ($a % 5)->sum + pdl(0) });
  both("ipow $t", sub { ($a % 3)->ipow(3) }) if !$t->integer || $t == longlong;
  if (!$t->integer) {
    { # libm vs CUDA math library: transcendentals agree within a few ulp, not bit for bit
      my $arg = (abs($a) % 9) + 1;
      PDL::B200::enable(1); my $g = $arg->log10; PDL::B200::enable(0); my $c = $arg->log10; PDL::B200::enable(1);
      ok(all(abs($g - $c) <= 4e-7 * abs($c) + 1e-30), "log10 $t within tolerance of the CPU path");
    }
    my $sp = $a->copy; $sp->set(3, 1, 'nan'); $sp->set(4, 2, 'inf');
    both("setnantobad $t",       sub { $sp->setnantobad });
    both("setnantobad clean $t", sub { $a->setnantobad });
    both("setinftobad $t",       sub { $sp->setinftobad });
    both("setnonfinitetobad $t", sub { $sp->setnonfinitetobad });
    both("setbadtonan $t",       sub { $bad->setbadtonan });
    both("isnan $t",             sub { $sp->isnan });
  }
}
# complex float / double: plus minus mult divide run on the device (gcc's inline multiply, libgcc's division restated)
for my $t (cfloat, cdouble) {
  my $rt = $t == cfloat ? float : double;
  my $za = PDL::czip(mk($rt, 300, 7), mk($rt, 300, 7)); my $zb = PDL::czip(mk($rt, 300, 7), mk($rt, 300, 7) + 0.5);
  both("complex plus $t",   sub { $za + $zb });
  both("complex minus $t",  sub { $za - $zb });
  both("complex mult $t",   sub { $za * $zb });
  both("complex divide $t", sub { $za / $zb });
  both("complex times real scalar $t", sub { $za * 2.5 });
  both("complex broadcast $t", sub { $za / $zb->slice(':,(3)') });
}
# large: exercises the device store (outputs created in it, inputs adopted on first use)
{
  my $y = sequence(2048, 2048); my $c = sequence(2048, 2048) * 0.5 + 1;
  both("cfg1 2048x2048 double", sub { $y + $c });
  both("chained, stays on device", sub { (($y + $c) * 2 - $y)->sumover });   # every value exactly representable
  my $x = $y + $c;
  is(PDL::B200::store_state($x), 4, 'op output lives in the device store (device copy current, host mirror not populated)');
  ok(PDL::B200::store_state($y) >= 4, 'large input was adopted into the device store');
  is($x->at(5, 7), $y->at(5, 7) + $c->at(5, 7), 'host read of a device result (lazy sync)');
  $x->set(0, 0, 42); my $z = $x + 1;
  is($z->at(0, 0), 43, 'host write then device op sees the new value');
  my $v = $x->slice('1:-1:2,3:9'); $v += 1000;
  is($x->at(1, 3), $y->at(1, 3) + $c->at(1, 3) + 1000, 'inplace through a slice reaches the parent');
  my ($ma, $mb) = (sequence(200, 300) % 64 - 32, (sequence(100, 200) % 32 - 16) / 16);  # exact products and sums
  both("matmult 300x200 double (tensor-core path, exact inputs)", sub { $ma x $mb });
}
my @st = PDL::B200::stats();
ok($st[0] > 100, "device readdata calls: $st[0], host calls: $st[1], migrated: $st[2], staged: $st[3], kernels: $st[4]");
done_testing;
