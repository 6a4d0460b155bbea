#!/usr/bin/env perl
# Generates the golden fixtures under tests/golden/*.json by running the REAL reference
# (PDL 2.106 built into oracle/_ref by oracle/build_ref.sh).  Committed together with its
# output so the fixtures can be regenerated; it is never run on the GPU box.
#
#   perl -Ioracle/_ref/blib/lib -Ioracle/_ref/blib/arch tests/golden/make_golden.pl tests/golden
#
# Every case records the INPUT bytes as the reference held them, the view chain applied to
# each input (slice/dummy/xchg/mv strings), the call, and the reference's OUTPUT (type, dims,
# badflag, bytes).  tests/test_golden.py replays the same call through pdl_b200's host logic
# on (a) the C oracle and (b) the CUDA path and compares.
use strict; use warnings;
use PDL::LiteF;
use JSON::PP;
PDL::set_autopthread_targ(0);

my $outdir = shift // 'tests/golden';
my @TYPES = qw(sbyte byte short ushort long ulong indx ulonglong longlong float double);
my %TOBJ = map { ($_ => PDL::Type->new($_)) } @TYPES;
my %IS_INT = map { ($_ => 1) } qw(sbyte byte short ushort long ulong indx ulonglong longlong);
my %IS_UNS = map { ($_ => 1) } qw(byte ushort ulong ulonglong);
my %BITS = (sbyte=>8, byte=>8, short=>16, ushort=>16, long=>32, ulong=>32, indx=>64, ulonglong=>64, longlong=>64);
my $INF = 9**9**9; my $NAN = -sin($INF);

my $seed = 12345;
sub rnd { $seed = ($seed * 1103515245 + 12345) % 2147483648; return $seed; }

sub hexdata { my ($p) = @_; my $q = $p->copy; $q->make_physical; return unpack('H*', ${$q->get_dataref}); }

# values for a type: flavour 'small' (|v| <= 9, good for products/shifts), 'mixed' (wide range,
# wraps on + - *), 'special' (floats: inf/nan/+-0 mixed in), 'pos' (strictly positive)
sub values_for {
  my ($type, $n, $flavour) = @_;
  my @v;
  for my $i (0..$n-1) {
    my $r = rnd();
    my $x;
    if ($IS_INT{$type}) {
      my $bits = $BITS{$type};
      if ($flavour eq 'small') { $x = $r % 19 - 9; }
      elsif ($flavour eq 'pos') { $x = $r % 9 + 1; }
      else {
        my $span = $bits >= 32 ? 2**31 - 1 : 2**($bits - ($IS_UNS{$type} ? 0 : 1)) - 1;
        $x = $r % (2 * $span + 1) - $span;
        $x = $span - ($r % 3) if $i % 7 == 3;          # hug the top of the range: forces wrap-around
      }
      $x = abs($x) if $IS_UNS{$type};
    } else {
      if ($flavour eq 'small') { $x = ($r % 19 - 9) / 2; }
      elsif ($flavour eq 'pos') { $x = ($r % 1000 + 1) / 8; }
      else { $x = (($r % 2000001) - 1000000) / 1024; }
      if ($flavour eq 'special') {
        my $k = $i % 11;
        $x = $INF if $k == 2; $x = -$INF if $k == 5; $x = $NAN if $k == 7; $x = 0 if $k == 8; $x = -0.0 if $k == 9;
      }
    }
    push @v, $x;
  }
  return @v;
}

sub mk { # ndarray of type/dims filled from values_for
  my ($type, $dims, $flavour) = @_;
  my $n = 1; $n *= $_ for @$dims;
  my $p = $n ? pdl($TOBJ{$type}, [values_for($type, $n, $flavour)]) : zeroes($TOBJ{$type}, 0);
  $p = $p->reshape(@$dims) if @$dims != 1 || $n == 0;
  return $p;
}

sub spec_of { # JSON description of a base (physical) input
  my ($p, $views) = @_;
  my $bv = $p->badvalue; $bv = pdl($p->type, $bv) unless ref $bv;
  return { type => $p->type->ioname, dims => [$p->dims], hex => hexdata($p), badflag => $p->badflag ? 1 : 0,
           badvalue_hex => unpack('H*', ${$bv->convert($p->type)->copy->get_dataref}),
           views => $views // [] };
}
sub apply_views {
  my ($p, $views) = @_;
  for my $v (@{$views // []}) {
    my ($m, @a) = @$v;
    $p = $p->$m(@a);
  }
  return $p;
}
sub out_of {
  my ($o) = @_;
  return { type => $o->type->ioname, dims => [$o->dims], hex => hexdata($o), badflag => $o->badflag ? 1 : 0 };
}

my @cases;
sub add_case {
  my ($name, $inputs, $call, $tol) = @_;   # inputs: [ [pdl, views] | {scalar=>..} ]
  my (@args, @specs);
  for my $in (@$inputs) {
    if (ref $in eq 'HASH') { push @args, $in->{scalar}; push @specs, { scalar => $in->{scalar}, is_int => ($in->{is_int} ? 1 : 0) }; }
    else { push @specs, spec_of($in->[0], $in->[1]); push @args, apply_views($in->[0], $in->[1]); }
  }
  my $out;
  my $kind = $call->{kind};
  my $ok = eval {
    if ($kind eq 'biop') {
      my $f = PDL->can($call->{op}) or die "no op $call->{op}";
      if ($call->{inplace}) { my $a = $args[0]->copy; $a->inplace; $out = $f->($a, $args[1], $call->{swap} // 0); }
      else { $out = $f->($args[0], $args[1], $call->{swap} // 0); }
    } elsif ($kind eq 'ufunc') {
      my $f = PDL->can($call->{op}) or die "no op $call->{op}";
      $out = $f->($args[0]);
    } elsif ($kind eq 'reduce' or $kind eq 'whole') {
      my $m = $call->{op};
      $out = $args[0]->$m;
    } elsif ($kind eq 'matmult') {
      $out = $args[0] x $args[1];
    } elsif ($kind eq 'ipow') {
      $out = PDL::ipow($args[0], $args[1]);
    } elsif ($kind eq 'convert') {
      $out = $args[0]->convert($TOBJ{$call->{to}});
    } elsif ($kind eq 'badop') {      # PDL::Bad functions: ndarray and scalar arguments in call order
      my $f = PDL->can($call->{op}) or die "no op $call->{op}";
      if ($call->{inplace}) { my $a = $args[0]->copy; $a->inplace; $f->($a, @args[1..$#args]); $out = $a; }
      else { $out = $f->(@args); }
    } elsif ($kind eq 'axis') {       # xvals / yvals / zvals of an existing ndarray
      my $m = $call->{op};
      $out = $args[0]->$m;
    } elsif ($kind eq 'sequence') {
      $out = defined $call->{type} ? sequence($TOBJ{$call->{type}}, @{$call->{dims}}) : sequence(@{$call->{dims}});
    } elsif ($kind eq 'inner') {
      $out = PDL::inner($args[0], $args[1]);
    } elsif ($kind eq 'outer') {
      $out = PDL::outer($args[0], $args[1]);
    } elsif ($kind eq 'minmaximum') {   # four outputs
      $out = [ $args[0]->minmaximum ];
    } elsif ($kind eq 'n_ind') {        # minimum_n_ind / maximum_n_ind with the size given
      my $m = $call->{op};
      $out = $args[0]->$m($call->{m});
    } else { die "kind $kind" }
    1;
  };
  if (!$ok) { my $e = $@; $e =~ s/\s+at \S+ line \d+.*//s; push @cases, { name => $name, inputs => \@specs, call => $call, error => $e }; return; }
  if (ref $out eq 'ARRAY') { push @cases, { name => $name, inputs => \@specs, call => $call, outputs => [map out_of($_), @$out] }; return; }
  push @cases, { name => $name, inputs => \@specs, call => $call, output => out_of($out), (defined $tol ? (tol_ulp => $tol) : ()) };
}
sub flush_cases {
  my ($file) = @_;
  open my $fh, '>', "$outdir/$file" or die "$outdir/$file: $!";
  print $fh JSON::PP->new->canonical->allow_nonref->encode({ reference => "PDL $PDL::VERSION", cases => \@cases });
  close $fh;
  printf "%-28s %4d cases\n", $file, scalar @cases;
  @cases = ();
}

# ---------------------------------------------------------------- biops (Ops.pd:288-313)
for my $t (@TYPES) {
  for my $op (qw(plus minus mult)) {
    add_case("$op-$t-mixed", [[mk($t,[7,3],'mixed')], [mk($t,[7,3],'mixed')]], {kind=>'biop', op=>$op});
    add_case("$op-$t-swap-scalar", [[mk($t,[5],'small')], {scalar=>3, is_int=>1}], {kind=>'biop', op=>$op, swap=>1});
  }
  add_case("divide-$t", [[mk($t,[9,2],'mixed')], [mk($t,[9,2],'pos')]], {kind=>'biop', op=>'divide'});
  add_case("divide-$t-swap", [[mk($t,[6],'pos')], [mk($t,[6],'small')]], {kind=>'biop', op=>'divide', swap=>1});
  for my $op (qw(gt lt le ge eq ne)) {
    add_case("$op-$t", [[mk($t,[11],'small')], [mk($t,[11],'small')]], {kind=>'biop', op=>$op});
  }
  if ($IS_INT{$t}) {
    my $cnt = pdl($TOBJ{$t}, [map { $_ % ($BITS{$t} < 32 ? 8 : $BITS{$t} - 1) } 0..12]);
    my $src = $IS_UNS{$t} ? mk($t,[13],'mixed') : mk($t,[13],'small');
    add_case("shiftleft-$t", [[$src], [$cnt]], {kind=>'biop', op=>'shiftleft'});
    add_case("shiftright-$t", [[mk($t,[13],'mixed')], [$cnt]], {kind=>'biop', op=>'shiftright'});
    for my $op (qw(or2 and2 xor)) { add_case("$op-$t", [[mk($t,[10],'mixed')], [mk($t,[10],'mixed')]], {kind=>'biop', op=>$op}); }
    add_case("bitnot-$t", [[mk($t,[10],'mixed')]], {kind=>'ufunc', op=>'bitnot'});
  } else {
    for my $op (qw(plus minus mult divide gt lt eq ne)) {
      add_case("$op-$t-special", [[mk($t,[22],'special')], [mk($t,[22],'special')]], {kind=>'biop', op=>$op});
    }
  }
  add_case("plus-$t-inplace", [[mk($t,[6,2],'small')], [mk($t,[6],'small')]], {kind=>'biop', op=>'plus', inplace=>1});
}
flush_cases('biop.json');

# ---------------------------------------------------------------- bifunc (Ops.pd:321-324)
for my $t (@TYPES) {
  my $b = mk($t,[17],'small'); # contains zeros and (for signed) negatives: MOD's branches
  add_case("modulo-$t", [[mk($t,[17],'mixed')], [$b]], {kind=>'biop', op=>'modulo'});
  add_case("modulo-$t-scalar", [[mk($t,[9],'mixed')], {scalar=>5, is_int=>1}], {kind=>'biop', op=>'modulo'});
  add_case("spaceship-$t", [[mk($t,[15],'small')], [mk($t,[15],'small')]], {kind=>'biop', op=>'spaceship'});
}
for my $t (qw(float double)) {
  add_case("power-$t", [[mk($t,[12],'pos')], [mk($t,[12],'small')]], {kind=>'biop', op=>'power'}, 4);
  add_case("atan2-$t", [[mk($t,[12],'small')], [mk($t,[12],'small')]], {kind=>'biop', op=>'atan2'}, 4);
  add_case("spaceship-$t-special", [[mk($t,[22],'special')], [mk($t,[22],'small')]], {kind=>'biop', op=>'spaceship'});
}
add_case("power-long-to-double", [[mk('long',[6],'pos')], {scalar=>2, is_int=>1}], {kind=>'biop', op=>'power'}, 4);
flush_cases('bifunc.json');

# ---------------------------------------------------------------- ufunc (Ops.pd:327-397,491-503)
for my $t (@TYPES) {
  add_case("sqrt-$t", [[mk($t,[10],'pos')]], {kind=>'ufunc', op=>'sqrt'}, $IS_INT{$t} ? undef : 0);
  add_case("sin-$t", [[mk($t,[10],'small')]], {kind=>'ufunc', op=>'sin'}, $IS_INT{$t} ? undef : 4);
  add_case("cos-$t", [[mk($t,[10],'small')]], {kind=>'ufunc', op=>'cos'}, $IS_INT{$t} ? undef : 4);
  add_case("not-$t", [[mk($t,[10],'small')]], {kind=>'ufunc', op=>'not'});
  add_case("log10-$t", [[mk($t,[10],'pos')]], {kind=>'ufunc', op=>'log10'}, $IS_INT{$t} ? undef : 4);
  add_case("abs-$t", [[mk($t,[12],'mixed')]], {kind=>'ufunc', op=>'abs'});
  add_case("abs2-$t", [[mk($t,[12],'small')]], {kind=>'ufunc', op=>'abs2'});
}
for my $t (qw(float double)) {
  add_case("exp-$t", [[mk($t,[12],'small')]], {kind=>'ufunc', op=>'exp'}, 4);
  add_case("log-$t", [[mk($t,[12],'pos')]], {kind=>'ufunc', op=>'log'}, 4);
  add_case("sqrt-$t-special", [[mk($t,[22],'special')]], {kind=>'ufunc', op=>'sqrt'}, 0);
  add_case("abs-$t-special", [[mk($t,[22],'special')]], {kind=>'ufunc', op=>'abs'});
}
add_case("exp-long-to-double", [[mk('long',[6],'small')]], {kind=>'ufunc', op=>'exp'}, 4);
# ipow (Ops.pd:443-476): a(); longlong b(); [o]ans()
for my $t (qw(ulonglong longlong float double long)) {
  my $base = mk($t,[12],'small'); $base->where($base == 0) .= 2 if $t =~ /long$/ ;   # 1/0 for negative powers kills the reference
  my $e = pdl(longlong, [0,1,2,3,5,7,8,-1,-2,13,-3,4]);
  add_case("ipow-$t", [[$base],[$e]], {kind=>'ipow'}, undef);
  add_case("ipow-$t-scalar", [[mk($t,[7],'pos')], {scalar=>5, is_int=>1}], {kind=>'ipow'}, undef);
}
flush_cases('ufunc.json');

# ---------------------------------------------------------------- type coercion + conversion (pdlapi.c:1182-1311, pdlconv.c:45-126)
add_case("coerce-byte+short", [[mk('byte',[5],'mixed')], [mk('short',[5],'mixed')]], {kind=>'biop', op=>'plus'});
add_case("coerce-float+double-scalar", [[mk('float',[5],'mixed')], {scalar=>1.5}], {kind=>'biop', op=>'plus'});
add_case("coerce-float+int-scalar", [[mk('float',[5],'mixed')], {scalar=>1, is_int=>1}], {kind=>'biop', op=>'plus'});
add_case("coerce-byte+300", [[mk('byte',[5],'mixed')], {scalar=>300, is_int=>1}], {kind=>'biop', op=>'plus'});
add_case("coerce-byte+100000", [[mk('byte',[5],'mixed')], {scalar=>100000, is_int=>1}], {kind=>'biop', op=>'mult'});
add_case("coerce-long*double", [[mk('long',[5],'mixed')], [mk('double',[5],'mixed')]], {kind=>'biop', op=>'mult'});
add_case("coerce-ulonglong-longlong", [[mk('ulonglong',[5],'mixed')], [mk('longlong',[5],'mixed')]], {kind=>'biop', op=>'minus'});
add_case("coerce-float-shift", [[mk('float',[5],'pos')], {scalar=>1, is_int=>1}], {kind=>'biop', op=>'shiftleft'});
add_case("coerce-ushort-gt-sbyte", [[mk('ushort',[5],'mixed')], [mk('sbyte',[5],'mixed')]], {kind=>'biop', op=>'gt'});
for my $from (@TYPES) { for my $to (@TYPES) {
  next if $from eq $to;
  my $fl = ($IS_UNS{$to} || $IS_UNS{$from}) ? 'pos' : 'small';
  add_case("convert-$from-$to", [[mk($from,[9],$fl)]], {kind=>'convert', to=>$to});
}}
flush_cases('coerce.json');

# ---------------------------------------------------------------- broadcasting over views (pdlbroadcast.c:275-488)
{
  my $big1 = mk('double',[16],'mixed'); my $big2 = mk('double',[16],'mixed');
  add_case("outer-dummy-slices", [[$big1, [['slice','0:-1:2'],['dummy',1,1]]], [$big2, [['slice','0:-1:2'],['dummy',0,1]]]], {kind=>'biop', op=>'mult'});
  my $m = mk('long',[6,5],'mixed');
  add_case("xchg-plus", [[$m, [['xchg',0,1]]], [mk('long',[5,6],'mixed')]], {kind=>'biop', op=>'plus'});
  add_case("reversed-slice", [[$m, [['slice','-1:0,:']]], [$m]], {kind=>'biop', op=>'minus'});
  add_case("row-broadcast", [[$m], [mk('long',[6],'small')]], {kind=>'biop', op=>'mult'});
  add_case("col-broadcast", [[$m], [mk('long',[1,5],'small')]], {kind=>'biop', op=>'mult'});
  add_case("3d-broadcast", [[mk('float',[4,1,3],'mixed')], [mk('float',[1,5,1],'mixed')]], {kind=>'biop', op=>'plus'});
  add_case("4d-mixed-views", [[mk('short',[3,4,2,2],'mixed'), [['mv',0,2]]], [mk('short',[4,2],'small')]], {kind=>'biop', op=>'plus'});
  add_case("strided-sub", [[mk('double',[20,4],'mixed'), [['slice','1:18:3,1:2']]], [mk('double',[6],'small')]], {kind=>'biop', op=>'plus'});
  add_case("dummy-big", [[mk('byte',[5],'mixed'), [['dummy',0,37]]], {scalar=>1, is_int=>1}], {kind=>'biop', op=>'plus'});
  add_case("empty-dim", [[mk('double',[0,3],'mixed')], [mk('double',[3],'mixed'), [['dummy',0,1]]]], {kind=>'biop', op=>'plus'});
  add_case("mismatch-error", [[mk('double',[3],'mixed')], [mk('double',[4],'mixed')]], {kind=>'biop', op=>'plus'});
  add_case("sqrt-of-slice", [[mk('double',[30],'pos'), [['slice','29:0:-3']]]], {kind=>'ufunc', op=>'sqrt'}, 0);
}
flush_cases('broadcast.json');

# ---------------------------------------------------------------- BAD values (Ops.pd:142-148,210,258; pdlapi.c:760-808)
sub with_bad { my ($p, @at) = @_; $p = $p->copy; $p->badflag(1); my $f = $p->flat; $f->setbadat($_) for @at; $p }
for my $t (@TYPES) {
  my $a = with_bad(mk($t,[12],'small'), 1, 5, 11); my $b = with_bad(mk($t,[12],'pos'), 0, 5);
  for my $op (qw(plus mult divide gt eq)) { add_case("bad-$op-$t", [[$a],[$b]], {kind=>'biop', op=>$op}); }
  add_case("bad-plus-$t-onesided", [[$a],[mk($t,[12],'small')]], {kind=>'biop', op=>'plus'});
  add_case("bad-modulo-$t", [[$a],[$b]], {kind=>'biop', op=>'modulo'});
  add_case("bad-abs-$t", [[$a]], {kind=>'ufunc', op=>'abs'});
  add_case("bad-not-$t", [[$a]], {kind=>'ufunc', op=>'not'});
  add_case("bad-inplace-$t", [[$a],[$b]], {kind=>'biop', op=>'plus', inplace=>1});
}
{
  # badflag set, data equal to the badvalue in the OTHER operand that has no badflag: biop checks the state flag
  my $a = with_bad(mk('long',[6],'small'), 2); my $b = pdl(long, [1, -2147483648, 3, 4, -2147483648, 6]);
  add_case("bad-state-check-biop", [[$a],[$b]], {kind=>'biop', op=>'plus'});
  add_case("bad-state-check-bifunc", [[$a],[$b]], {kind=>'biop', op=>'spaceship'});
  # per-ndarray badvalue
  my $c = mk('short',[8],'small')->copy; $c->badflag(1); $c->badvalue(3);
  add_case("bad-custom-badvalue", [[$c],[mk('short',[8],'small')]], {kind=>'biop', op=>'plus'});
  # NaN as badvalue
  for my $t (qw(float double)) {
    my $d = mk($t,[22],'special')->copy; $d->badflag(1); $d->badvalue($NAN);
    add_case("bad-nan-badvalue-$t", [[$d],[mk($t,[22],'small')]], {kind=>'biop', op=>'plus'});
    add_case("bad-nan-badvalue-sumover-$t", [[$d]], {kind=>'reduce', op=>'sumover'});
    add_case("bad-nan-badvalue-minimum-$t", [[$d]], {kind=>'reduce', op=>'minimum'});
  }
  add_case("bad-through-slice", [[with_bad(mk('double',[10,3],'mixed'), 3, 14, 27), [['slice','1:-1:2,:']]], {scalar=>2.5}], {kind=>'biop', op=>'mult'});
  add_case("bad-convert", [[with_bad(mk('float',[9],'small'), 1, 4)]], {kind=>'convert', to=>'short'});
  add_case("bad-coerce-plus", [[with_bad(mk('byte',[9],'small'), 1, 4)], [mk('float',[9],'small')]], {kind=>'biop', op=>'plus'});
}
flush_cases('bad.json');

# ---------------------------------------------------------------- reductions (Ufunc.pd:88-187,413-500)
my @RED = qw(sumover prodover dsumover dprodover average daverage minimum maximum minimum_ind maximum_ind
             andover orover zcover xorover);
for my $t (@TYPES) {
  my $a = mk($t,[13,4],'small');
  for my $op (@RED, ($IS_INT{$t} ? qw(bandover borover bxorover) : ())) {
    my $tol = (!$IS_INT{$t} && $op =~ /^(d?sumover|d?prodover|d?average)$/) ? 0 : undef;  # exactly representable inputs
    add_case("$op-$t", [[$a]], {kind=>'reduce', op=>$op}, $tol);
  }
  add_case("sumover-$t-wrap", [[mk($t,[40,2],'mixed')]], {kind=>'reduce', op=>'sumover'});
  add_case("sumover-$t-xchg", [[mk($t,[5,7],'small'), [['xchg',0,1]]]], {kind=>'reduce', op=>'sumover'}, $IS_INT{$t} ? undef : 0);
  add_case("maximum_ind-$t-strided", [[mk($t,[24,3],'small'), [['slice','-1:0:-2,:']]]], {kind=>'reduce', op=>'maximum_ind'});
  add_case("minimum-$t-3d", [[mk($t,[6,3,4],'mixed')]], {kind=>'reduce', op=>'minimum'});
  my $b = with_bad(mk($t,[9,4],'small'), 0, 3, 8, (map { 18 + $_ } 0..8), 30);   # row 2 is all BAD
  for my $op (qw(sumover prodover average daverage minimum maximum minimum_ind maximum_ind andover orover)) {
    add_case("bad-$op-$t", [[$b]], {kind=>'reduce', op=>$op}, (!$IS_INT{$t} && $op =~ /sum|prod|aver/) ? 0 : undef);
  }
  add_case("empty-n-sumover-$t", [[mk($t,[0,3],'small')]], {kind=>'reduce', op=>'sumover'});
  add_case("empty-n-maximum-$t", [[mk($t,[0,3],'small')]], {kind=>'reduce', op=>'maximum'});
  add_case("empty-n-average-$t", [[mk($t,[0,2],'small')]], {kind=>'reduce', op=>'average'});
  add_case("whole-sum-$t", [[mk($t,[7,5],'small')]], {kind=>'whole', op=>'sum'}, $IS_INT{$t} ? undef : 0);
  add_case("whole-max-$t", [[mk($t,[7,5],'mixed')]], {kind=>'whole', op=>'max'});
  add_case("whole-min-$t-1d", [[mk($t,[50],'mixed')]], {kind=>'whole', op=>'min'});
}
for my $t (qw(float double)) {
  # NaN handling and signed zeros in min/max (Ufunc.pd:460; t/ufunc.t:82-98)
  my $rows = pdl($TOBJ{$t}, [[$NAN,1,2],[1,$NAN,2],[1,2,$NAN],[$NAN,$NAN,$NAN],[0,-0.0,0],[-0.0,0,-0.0],[$INF,-$INF,5],[3,3,3]]);
  for my $op (qw(minimum maximum minimum_ind maximum_ind sumover average prodover)) {
    add_case("nan-zero-$op-$t", [[$rows]], {kind=>'reduce', op=>$op});
  }
  add_case("prodover-zero-inf-$t", [[pdl($TOBJ{$t}, [[0,$INF,2],[$INF,0,2],[2,0,-3],[-2,0,3]])]], {kind=>'reduce', op=>'prodover'});
  add_case("long-row-sumover-$t", [[mk($t,[5000,3],'small')]], {kind=>'reduce', op=>'sumover'}, 0);
  add_case("long-row-average-$t", [[mk($t,[5000,3],'small')]], {kind=>'reduce', op=>'average'}, 0);
}
# counts (Bad.pd:418-480) and scans (Ufunc.pd:120-141)
for my $t (@TYPES) {
  my $b = with_bad(mk($t,[11,3],'small'), 0, 5, 12, (map { 22 + $_ } 0..10));   # row 2 all BAD
  for my $op (qw(nbadover ngoodover)) {
    add_case("$op-$t", [[$b]], {kind=>'reduce', op=>$op});
    add_case("$op-$t-noflag", [[mk($t,[7,2],'small')]], {kind=>'reduce', op=>$op});
  }
  for my $op (qw(cumusumover cumuprodover dcumusumover dcumuprodover)) {
    add_case("$op-$t", [[mk($t,[9,3],'small')]], {kind=>'reduce', op=>$op}, $IS_INT{$t} ? undef : 0);
    add_case("bad-$op-$t", [[$b]], {kind=>'reduce', op=>$op}, $IS_INT{$t} ? undef : 0);
  }
  add_case("cumusumover-$t-wrap-xchg", [[mk($t,[6,20],'mixed'), [['xchg',0,1]]]], {kind=>'reduce', op=>'cumusumover'});
}
flush_cases('reduce.json');

# ---------------------------------------------------------------- matmult (Primitive.pd:191-264; t/primitive-matmult.t)
{
  my $pa = pdl([[1,2,3,4,5],[6,7,8,9,10],[11,12,13,14,15],[16,17,18,19,20],[21,22,23,24,25]]);
  add_case("matmult-5x5-fiducial", [[$pa],[$pa]], {kind=>'matmult'});
  my $pb = pdl([[1,2,3,4],[5,6,7,8],[9,10,11,12]]); my $pc = pdl([[1,2],[3,4],[5,6],[7,8]]);
  add_case("matmult-3x4-4x2", [[$pb],[$pc]], {kind=>'matmult'});
  add_case("matmult-sliced-dummy", [[sequence(5,3)],[sequence(5), [['dummy',0,1]]]], {kind=>'matmult'});  # t/primitive-matmult.t:53-57 shape
  add_case("matmult-vector", [[sequence(3,2)],[sequence(3), [['dummy',0,1]]]], {kind=>'matmult'});
  add_case("matmult-dim-mismatch", [[sequence(3,2)],[sequence(2,2)]], {kind=>'matmult'});
  add_case("matmult-scalar-shortcut", [[sequence(3,2)],[pdl(2)]], {kind=>'matmult'});
  for my $t (@TYPES) {
    add_case("matmult-$t-17x23x9", [[mk($t,[23,17],'small')],[mk($t,[9,23],'small')]], {kind=>'matmult'});
    add_case("matmult-$t-wrap", [[mk($t,[70,3],'mixed')],[mk($t,[4,70],'mixed')]], {kind=>'matmult'}, $IS_INT{$t} ? undef : 0);
  }
  for my $t (qw(float double)) {
    add_case("matmult-$t-bitexact-order", [[mk($t,[67,65],'mixed')],[mk($t,[66,67],'mixed')]], {kind=>'matmult'}, 0);
    add_case("matmult-$t-nan", [[pdl($TOBJ{$t}, [[1,$NAN],[3,4]])],[pdl($TOBJ{$t}, [[1,2],[3,4]])]], {kind=>'matmult'});
    my $ab = with_bad(mk($t,[5,4],'small'), 7); my $bb = with_bad(mk($t,[3,5],'small'), 4);
    add_case("matmult-$t-bad-a", [[$ab],[mk($t,[3,5],'small')]], {kind=>'matmult'});
    add_case("matmult-$t-bad-both", [[$ab],[$bb]], {kind=>'matmult'});
  }
  add_case("matmult-long-bad", [[with_bad(mk('long',[20,4],'small'), 7, 50)],[mk('long',[3,20],'small')]], {kind=>'matmult'});
  add_case("matmult-batched", [[mk('double',[4,3,5],'small')],[mk('double',[2,4,5],'small')]], {kind=>'matmult'});
  add_case("matmult-batched-broadcast-b", [[mk('double',[4,3,5],'small')],[mk('double',[2,4],'small')]], {kind=>'matmult'});
  add_case("matmult-transposed-view", [[mk('double',[6,8],'small'), [['xchg',0,1]]],[mk('double',[8,5],'small'), [['xchg',0,1]]]], {kind=>'matmult'});
}
flush_cases('matmult.json');

# ---------------------------------------------------------------- Bad.pd elementwise ops (Bad.pd:343-416,584-905)
for my $t (@TYPES) {
  my $a = with_bad(mk($t,[12],'small'), 1, 5, 11);
  my $g = mk($t,[12],'small');
  for my $op (qw(isbad isgood isnan)) {
    add_case("$op-$t-bad", [[$a]], {kind=>'badop', op=>$op});
    add_case("$op-$t-good", [[$g]], {kind=>'badop', op=>$op});
  }
  my $mask = pdl(long, [0,1,0,0,-3,0,0,0,2,0,0,0]);
  add_case("setbadif-$t", [[$g],[$mask]], {kind=>'badop', op=>'setbadif'});
  add_case("setbadif-$t-badin", [[$a],[with_bad($mask, 3)]], {kind=>'badop', op=>'setbadif'});
  add_case("setbadif-$t-2d-rowmask", [[mk($t,[5,3],'small')],[pdl(long,[1,0,0,1,0])]], {kind=>'badop', op=>'setbadif'});
  add_case("setbadif-$t-doublemask", [[$g],[pdl(double,[0,0.5,1,0,2.5,0,0,0,0,-1,0,0])]], {kind=>'badop', op=>'setbadif'});
  add_case("setvaltobad-$t", [[$g],{scalar=>3, is_int=>1}], {kind=>'badop', op=>'setvaltobad'});
  add_case("setvaltobad-$t-inplace", [[$a],{scalar=>2, is_int=>1}], {kind=>'badop', op=>'setvaltobad', inplace=>1});
  add_case("setbadtoval-$t", [[$a],{scalar=>7, is_int=>1}], {kind=>'badop', op=>'setbadtoval'});
  add_case("setbadtoval-$t-good", [[$g],{scalar=>7, is_int=>1}], {kind=>'badop', op=>'setbadtoval'});
  add_case("setbadtoval-$t-inplace", [[$a],{scalar=>1, is_int=>1}], {kind=>'badop', op=>'setbadtoval', inplace=>1});
  add_case("badmask-$t", [[$a],[mk($t,[12],'pos')]], {kind=>'badop', op=>'badmask'});
  add_case("badmask-$t-scalar", [[$a],{scalar=>4, is_int=>1}], {kind=>'badop', op=>'badmask'});
  add_case("copybad-$t", [[$g],[$a]], {kind=>'badop', op=>'copybad'});
  add_case("copybad-$t-good", [[$g],[mk($t,[12],'pos')]], {kind=>'badop', op=>'copybad'});
  add_case("setbadif-then-sumover-$t-view", [[mk($t,[20,3],'small'), [['slice','1:18:3,:']]],[pdl(long,[0,1,0,0,1,0])]], {kind=>'badop', op=>'setbadif'});
}
for my $t (qw(float double)) {
  my $sp = mk($t,[22],'special');
  my $fine = mk($t,[22],'small');
  for my $op (qw(setnantobad setinftobad setnonfinitetobad setbadtonan isnan)) {
    add_case("$op-$t-special", [[$sp]], {kind=>'badop', op=>$op});
    add_case("$op-$t-finite", [[$fine]], {kind=>'badop', op=>$op});
    add_case("$op-$t-badin", [[with_bad($sp, 0, 3)]], {kind=>'badop', op=>$op});
  }
  add_case("setnantobad-$t-inplace", [[$sp]], {kind=>'badop', op=>'setnantobad', inplace=>1});
  add_case("setbadtonan-$t-inplace", [[with_bad($fine, 2, 9)]], {kind=>'badop', op=>'setbadtonan', inplace=>1});
  add_case("badmask-$t-special", [[$sp],{scalar=>-1, is_int=>1}], {kind=>'badop', op=>'badmask'});
  add_case("setvaltobad-$t-frac", [[$fine],{scalar=>1.5}], {kind=>'badop', op=>'setvaltobad'});
  # a value equal to the default badvalue in an ndarray WITHOUT the badflag: setbadtonan still tests it (Bad.pd:795)
  my $raw = $t eq 'float' ? pdl(float, [1, -3.4028234663852886e38, 3]) : pdl(double, [1, -1.7976931348623157e308, 3]);
  add_case("setbadtonan-$t-noflag-badvalue", [[$raw]], {kind=>'badop', op=>'setbadtonan'});
  add_case("setbadtoval-$t-noflag-badvalue", [[$raw],{scalar=>0, is_int=>1}], {kind=>'badop', op=>'setbadtoval'});
  my $d = mk($t,[22],'special')->copy; $d->badflag(1); $d->badvalue($NAN);
  add_case("isbad-$t-nan-badvalue", [[$d]], {kind=>'badop', op=>'isbad'});
  add_case("setbadtoval-$t-nan-badvalue", [[$d],{scalar=>9, is_int=>1}], {kind=>'badop', op=>'setbadtoval'});
}
add_case("setnantobad-long", [[mk('long',[6],'small')]], {kind=>'badop', op=>'setnantobad'});
add_case("setvaltobad-byte-300", [[mk('byte',[9],'mixed')],{scalar=>300, is_int=>1}], {kind=>'badop', op=>'setvaltobad'});
flush_cases('badops.json');

# ---------------------------------------------------------------- constructors (Basic.pm:117-129,479-485; Primitive.pd:1468-1474)
for my $t (@TYPES) {
  add_case("sequence-$t", [], {kind=>'sequence', type=>$t, dims=>[7,3]});
  add_case("sequence-$t-long", [], {kind=>'sequence', type=>$t, dims=>[300]});
  for my $op (qw(xvals yvals zvals)) {
    add_case("$op-$t", [[mk($t,[5,3,2],'small')]], {kind=>'axis', op=>$op});
  }
  add_case("yvals-$t-1d", [[mk($t,[5],'small')]], {kind=>'axis', op=>'yvals'});
}
add_case("sequence-untyped", [], {kind=>'sequence', dims=>[4,2]});
add_case("sequence-empty", [], {kind=>'sequence', type=>'long', dims=>[0,3]});
add_case("xvals-of-view", [[mk('double',[10,4],'small'), [['slice','1:8:2,:'],['xchg',0,1]]]], {kind=>'axis', op=>'xvals'});
flush_cases('basic.json');

# ---------------------------------------------------------------- inner (Primitive.pd:48-70)
for my $t (@TYPES) {
  add_case("inner-$t", [[mk($t,[13,4],'small')],[mk($t,[13,4],'small')]], {kind=>'inner'}, $IS_INT{$t} ? undef : 0);
  add_case("inner-$t-broadcast", [[mk($t,[9,3],'small')],[mk($t,[9],'small')]], {kind=>'inner'}, $IS_INT{$t} ? undef : 0);
  add_case("inner-$t-long-row", [[mk($t,[3000],'small')],[mk($t,[3000],'small')]], {kind=>'inner'}, $IS_INT{$t} ? undef : 0) unless $BITS{$t} && $BITS{$t} < 32;
  add_case("inner-$t-bad", [[with_bad(mk($t,[9,4],'small'), 3, 20)],[with_bad(mk($t,[9,4],'small'), 30)]], {kind=>'inner'}, $IS_INT{$t} ? undef : 0);
  add_case("inner-$t-outer-views", [[mk($t,[8],'small'), [['dummy',1,1]]],[mk($t,[6],'small'), [['dummy',0,1]]]], {kind=>'inner'}, $IS_INT{$t} ? undef : 0);
  add_case("inner-$t-empty-n", [[mk($t,[0,3],'small')],[mk($t,[0,3],'small')]], {kind=>'inner'});
}
add_case("inner-mismatch", [[mk('double',[3],'small')],[mk('double',[4],'small')]], {kind=>'inner'});
flush_cases('inner.json');

# ---------------------------------------------------------------- minmaximum (Ufunc.pd:563-613), magnover (:1235-1256)
for my $t (@TYPES) {
  add_case("minmaximum-$t", [[mk($t,[13,4],'small')]], {kind=>'minmaximum'});
  add_case("minmaximum-$t-mixed-3d", [[mk($t,[6,3,4],'mixed')]], {kind=>'minmaximum'});
  add_case("minmaximum-$t-long-flat", [[mk($t,[3000],'mixed')]], {kind=>'minmaximum'});
  add_case("minmaximum-$t-strided", [[mk($t,[24,3],'small'), [['slice','-1:0:-2,:']]]], {kind=>'minmaximum'});
  add_case("minmaximum-$t-bad", [[with_bad(mk($t,[9,4],'small'), 0, 3, 8, (map { 18 + $_ } 0..8), 30)]], {kind=>'minmaximum'});
  add_case("minmaximum-$t-empty-n", [[mk($t,[0,3],'small')]], {kind=>'minmaximum'});
  add_case("minmaximum-$t-ties", [[pdl($TOBJ{$t}, [[3,1,1,3,2],[5,5,5,5,5]])]], {kind=>'minmaximum'});
}
for my $t (qw(float double)) {
  my $rows = pdl($TOBJ{$t}, [[$NAN,1,2],[1,$NAN,2],[1,2,$NAN],[$NAN,$NAN,$NAN],[0,-0.0,0],[-0.0,0,-0.0],[$INF,-$INF,5],[3,3,3]]);
  add_case("minmaximum-$t-nan-zero", [[$rows]], {kind=>'minmaximum'});
  my $rb = $rows->copy; $rb->badflag(1); $rb->badvalue($NAN);
  add_case("minmaximum-$t-nan-badvalue", [[$rb]], {kind=>'minmaximum'});
  add_case("magnover-$t", [[mk($t,[13,4],'small')]], {kind=>'reduce', op=>'magnover'}, 1);
  add_case("magnover-$t-345", [[pdl($TOBJ{$t}, [[3,4],[0,0],[5,12],[-8,15]])]], {kind=>'reduce', op=>'magnover'}, 0);
  add_case("magnover-$t-bad", [[with_bad(mk($t,[9,4],'small'), 0, 3, 8, (map { 18 + $_ } 0..8), 30)]], {kind=>'reduce', op=>'magnover'}, 1);
  add_case("magnover-$t-long", [[mk($t,[5000,2],'small')]], {kind=>'reduce', op=>'magnover'}, 1);
  add_case("magnover-$t-empty-n", [[mk($t,[0,3],'small')]], {kind=>'reduce', op=>'magnover'});
}
add_case("magnover-long-to-float", [[mk('long',[7,2],'small')]], {kind=>'reduce', op=>'magnover'}, 1);
flush_cases('minmax.json');

# ---------------------------------------------------------------- outer (Primitive.pd:78-96)
for my $t (@TYPES) {
  add_case("outer-$t", [[mk($t,[13],'small')],[mk($t,[7],'small')]], {kind=>'outer'});
  add_case("outer-$t-broadcast", [[mk($t,[5,3],'small')],[mk($t,[4],'mixed')]], {kind=>'outer'});
  add_case("outer-$t-bad", [[with_bad(mk($t,[9],'small'), 3)],[with_bad(mk($t,[6,2],'small'), 7)]], {kind=>'outer'});
  add_case("outer-$t-strided", [[mk($t,[24],'small'), [['slice','-1:0:-2']]],[mk($t,[5,4],'small'), [['xchg',0,1]]]], {kind=>'outer'});
}
add_case("outer-empty", [[mk('double',[0],'small')],[mk('double',[3],'small')]], {kind=>'outer'});
flush_cases('outer.json');

# ---------------------------------------------------------------- edge shapes for the rows either side of the path
{
  my $m3 = mk('double',[6,4,3],'small');
  add_case("edge-setbadif-3d-broadcast-mask", [[$m3],[pdl(long,[[0,1,0,0,1,0]])]], {kind=>'badop', op=>'setbadif'});
  add_case("edge-setbadif-col-mask", [[mk('short',[5,4],'small')],[pdl(long,[1,0,0,1])->dummy(0,1)]], {kind=>'badop', op=>'setbadif'});
  add_case("edge-setbadtoval-reversed-view", [[with_bad(mk('long',[12],'small'), 2, 7), [['slice','-1:0']]],{scalar=>-5, is_int=>1}], {kind=>'badop', op=>'setbadtoval'});
  add_case("edge-isbad-0dim", [[with_bad(mk('float',[1],'small'), 0), [['slice','(0)']]]], {kind=>'badop', op=>'isbad'});
  add_case("edge-isgood-empty", [[mk('double',[0,3],'small')]], {kind=>'badop', op=>'isgood'});
  add_case("edge-copybad-broadcast", [[mk('float',[7,3],'small')],[with_bad(mk('float',[7],'small'), 1, 5)]], {kind=>'badop', op=>'copybad'});
  add_case("edge-badmask-nan-badvalue", [[do { my $d = mk('double',[22],'special')->copy; $d->badflag(1); $d->badvalue($NAN); $d }],{scalar=>0, is_int=>1}], {kind=>'badop', op=>'badmask'});
  add_case("edge-setvaltobad-negative-on-unsigned", [[mk('byte',[9],'mixed')],{scalar=>-1, is_int=>1}], {kind=>'badop', op=>'setvaltobad'});
  add_case("edge-setbadtoval-frac-on-int", [[with_bad(mk('long',[9],'small'), 1, 4)],{scalar=>2.75}], {kind=>'badop', op=>'setbadtoval'});
  add_case("edge-inner-reversed", [[mk('double',[40],'small'), [['slice','-1:0']]],[mk('double',[40],'small')]], {kind=>'inner'}, 0);
  add_case("edge-inner-stride3-x-dummy", [[mk('float',[30,4],'small'), [['slice','0:-1:3,:']]],[mk('float',[10],'small')]], {kind=>'inner'}, 0);
  add_case("edge-inner-xchg", [[mk('long',[6,9],'small'), [['xchg',0,1]]],[mk('long',[9],'small')]], {kind=>'inner'});
  add_case("edge-inner-scalar-b", [[mk('double',[5,3],'small')],{scalar=>2.5}], {kind=>'inner'}, 0);
  add_case("edge-outer-dummy", [[mk('long',[4],'small'), [['dummy',1,3]]],[mk('long',[5],'small')]], {kind=>'outer'});
  add_case("edge-outer-1x1", [[mk('double',[1],'small')],[mk('double',[1],'small')]], {kind=>'outer'});
  add_case("edge-minmaximum-1elem", [[mk('float',[1,5],'mixed')]], {kind=>'minmaximum'});
  add_case("edge-minmaximum-xchg", [[mk('short',[7,40],'small'), [['xchg',0,1]]]], {kind=>'minmaximum'});
  add_case("edge-minmaximum-inf", [[pdl(double, [[$INF,-$INF,1],[$INF,$INF,$INF],[-$INF,-$INF,-$INF],[$NAN,$INF,$NAN]])]], {kind=>'minmaximum'});
  add_case("edge-minmaximum-intmax", [[pdl(long, [[2147483647,2147483647],[-2147483648,-2147483648],[2147483647,-2147483648]])]], {kind=>'minmaximum'});
  add_case("edge-minmaximum-bytes-ident", [[pdl(byte, [[255,255,255],[0,0,0],[0,255,0]])]], {kind=>'minmaximum'});
  add_case("edge-sequence-1", [], {kind=>'sequence', type=>'double', dims=>[1]});
  add_case("edge-sequence-byte-wrap", [], {kind=>'sequence', type=>'byte', dims=>[300]});
  add_case("edge-sequence-float-big", [], {kind=>'sequence', type=>'float', dims=>[5,7,3,2]});
  add_case("edge-zvals-2d", [[mk('double',[4,3],'small')]], {kind=>'axis', op=>'zvals'});
  add_case("edge-yvals-reversed-view", [[mk('float',[6,5],'small'), [['slice',':,-1:0']]]], {kind=>'axis', op=>'yvals'});
  for my $t (qw(byte sbyte short ushort)) {
    my $row = mk($t,[700,3],'mixed');                     # rows long enough for the packed 16-byte paths, odd alignment via slice
    my $b = with_bad($row, 5, 699, 700, 1399, (map { 1400 + $_ } 0..699));
    for my $op (qw(sumover average minimum maximum orover andover bandover borover bxorover zcover xorover)) {
      add_case("edge-packed-$op-$t", [[$row, [['slice','3:-2,:']]]], {kind=>'reduce', op=>$op});
      add_case("edge-packed-bad-$op-$t", [[$b, [['slice','1:-1,:']]]], {kind=>'reduce', op=>$op});
    }
    add_case("edge-packed-minmaximum-$t", [[$b]], {kind=>'minmaximum'});
  }
  add_case("edge-magnover-xchg", [[mk('double',[5,60],'small'), [['xchg',0,1]]]], {kind=>'reduce', op=>'magnover'}, 1);
  add_case("edge-cumusumover-reversed", [[mk('long',[50,2],'small'), [['slice','-1:0,:']]]], {kind=>'reduce', op=>'cumusumover'});
  add_case("edge-cumuprodover-bad-first", [[with_bad(mk('double',[9,2],'small'), 0, 9)]], {kind=>'reduce', op=>'cumuprodover'}, 0);
}
flush_cases('edge.json');

# ---------------------------------------------------------------- round 2: minimum_n_ind / maximum_n_ind (Ufunc.pd:502-561),
# inner / magnover with non-finite values and overflow (Primitive.pd:48-70, Ufunc.pd:1235-1256)
for my $t (@TYPES) {
  for my $op (qw(minimum_n_ind maximum_n_ind)) {
    add_case("$op-$t", [[mk($t,[13,4],'small')]], {kind=>'n_ind', op=>$op, m=>3});
    add_case("$op-$t-all", [[mk($t,[6,2],'mixed')]], {kind=>'n_ind', op=>$op, m=>6});
    add_case("$op-$t-long", [[mk($t,[3000],'mixed')]], {kind=>'n_ind', op=>$op, m=>4});
    add_case("$op-$t-strided", [[mk($t,[24,3],'small'), [['slice','-1:0:-2,:']]]], {kind=>'n_ind', op=>$op, m=>5});
    # rows 1 and 2 cannot fill their slots, the LAST row can: the output ends up with BAD values and badflag 0
    add_case("$op-$t-bad-early-row", [[with_bad(mk($t,[5,4],'small'), 5..9, 10, 11, 12, 14)]], {kind=>'n_ind', op=>$op, m=>2});
    # the last row cannot: badflag 1
    add_case("$op-$t-bad-last-row", [[with_bad(mk($t,[5,4],'small'), 2, 15, 16, 17, 19)]], {kind=>'n_ind', op=>$op, m=>2});
    add_case("$op-$t-ties", [[pdl($TOBJ{$t}, [[3,1,1,3,2],[5,5,5,5,5]])]], {kind=>'n_ind', op=>$op, m=>4});
  }
}
for my $t (qw(float double)) {
  my $rows = pdl($TOBJ{$t}, [[$NAN,1,2,0],[1,$NAN,2,$NAN],[$NAN,$NAN,$NAN,$NAN],[0,-0.0,0,-0.0],[$INF,-$INF,5,$NAN]]);
  add_case("$_-$t-nan-zero", [[$rows]], {kind=>'n_ind', op=>$_, m=>3}) for qw(minimum_n_ind maximum_n_ind);
  my $big = $t eq 'float' ? 3e38 : 1e308;
  my $nf = pdl($TOBJ{$t}, [[1,$INF,2,3],[$big,$big,$big,1],[$INF,-$INF,1,1],[-$INF,5,1,1],[$NAN,1,2,3],[1,2,3,4]]);
  my $two = pdl($TOBJ{$t}, 2);
  add_case("inner-$t-nonfinite", [[$nf],[$two]], {kind=>'inner'}, 0);
  add_case("magnover-$t-nonfinite", [[$nf]], {kind=>'reduce', op=>'magnover'}, 0);
}
add_case("minimum_n_ind-too-many", [[mk('double',[3],'small')]], {kind=>'n_ind', op=>'minimum_n_ind', m=>4});
flush_cases('round2.json');

# ---------------------------------------------------------------- complex float / double: plus minus mult divide (Ops.pd:104-153,288-291)
$TOBJ{cfloat} = PDL::Type->new('cfloat'); $TOBJ{cdouble} = PDL::Type->new('cdouble');
for my $t (qw(cfloat cdouble)) {
  my $rt = $t eq 'cfloat' ? 'float' : 'double';
  my $mkc = sub { my ($dims, $fl) = @_; PDL::czip(mk($rt, $dims, $fl), mk($rt, $dims, $fl)) };
  my ($big, $tiny, $sub) = $t eq 'cfloat' ? (1e30, 1e-30, 1e-40) : (1e300, 1e-300, 1e-310);
  # magnitudes that reach every scaling branch of libgcc's __divdc3 (near overflow, near underflow, denormals, tiny ratios)
  my $ext = PDL::czip(pdl($TOBJ{$rt}, [$big, $tiny, $sub, 1, $big, $tiny, 3, -$big, $sub, 0, 1e10, $big]),
                      pdl($TOBJ{$rt}, [$big, $tiny, 1, $sub, $tiny, $big, -4, 2, $sub, 0, $tiny, -$big]));
  my $ext2 = PDL::czip(pdl($TOBJ{$rt}, [$big, $tiny, $tiny, $big, 1, 2, $sub, $big, 1, 1, $big, $tiny]),
                       pdl($TOBJ{$rt}, [$tiny, $big, $sub, $big, $sub, -2, 1, $big, $big, 0, 1e-10, $tiny]));
  my $spec = PDL::czip(mk($rt, [22], 'special'), mk($rt, [22], 'special'));
  my $spec2 = PDL::czip(mk($rt, [22], 'special'), mk($rt, [22], 'small'));
  for my $op (qw(plus minus mult divide)) {
    add_case("$op-$t", [[$mkc->([7,3],'mixed')], [$mkc->([7,3],'mixed')]], {kind=>'biop', op=>$op});
    add_case("$op-$t-small", [[$mkc->([50],'small')], [$mkc->([50],'pos')]], {kind=>'biop', op=>$op});
    add_case("$op-$t-broadcast", [[$mkc->([5,4],'small')], [$mkc->([5],'pos')]], {kind=>'biop', op=>$op});
    add_case("$op-$t-strided", [[$mkc->([12,3],'small'), [['slice','-1:0:-2,:']]], [$mkc->([6],'pos')]], {kind=>'biop', op=>$op});
    add_case("$op-$t-swap-real-scalar", [[$mkc->([6],'small')], {scalar=>2.5}], {kind=>'biop', op=>$op, swap=>1});
    add_case("$op-$t-extreme", [[$ext], [$ext2]], {kind=>'biop', op=>$op});
    add_case("$op-$t-extreme-swapped", [[$ext2], [$ext]], {kind=>'biop', op=>$op});
    add_case("$op-$t-special", [[$spec], [$spec2]], {kind=>'biop', op=>$op});
    add_case("$op-$t-special-swapped", [[$spec2], [$spec]], {kind=>'biop', op=>$op});
    add_case("$op-$t-bad", [[with_bad($mkc->([9,2],'small'), 1, 5, 11)], [with_bad($mkc->([9,2],'pos'), 0, 5)]], {kind=>'biop', op=>$op});
    add_case("$op-$t-inplace", [[$mkc->([8],'small')], [$mkc->([8],'pos')]], {kind=>'biop', op=>$op, inplace=>1});
  }
}
flush_cases('complex.json');
