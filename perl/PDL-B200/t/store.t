#!/usr/bin/env perl
# The device-resident data store under the UNMODIFIED core (perl/PDL-B200/pdlb200_pp.h, pdl_b200/csrc/store.cu):
# where ndarray data lives, when it crosses PCIe, and that no host-access path of the reference can see stale bytes.
use strict; use warnings;
use Test::More;
use PDL::LiteF;
use PDL::B200;

plan skip_all => 'no CUDA device' unless PDL::B200::device_count() > 0;
PDL::set_autopthread_targ(0);

sub st { my %h; @h{qw(dev host adopted staged kernels syncs transient cpu_trans)} = PDL::B200::stats(); \%h }
sub ss { my %h; @h{qw(new recycled uploads upload_bytes downloads download_bytes faults adopted)} = PDL::B200::store_stats(); \%h }
sub bytes { my $q = $_[0]->copy; unpack('H*', ${ $q->get_dataref }) }
sub cpu { my ($code) = @_; PDL::B200::enable(0); my $r = $code->(); PDL::B200::enable(1); $r }

my $y = sequence(2048, 2048) / 1024; my $c = sequence(2048, 2048) * 0.5 + 1;
{
  my ($s0, $m0) = (st(), ss());
  my $x = $y + $c;
  my ($s1, $m1) = (st(), ss());
  is(PDL::B200::store_state($x), 4, 'result: device copy current, host mirror untouched');
  is($s1->{syncs} - $s0->{syncs}, 0, 'no stream synchronisation after a device op on resident ndarrays');
  is($m1->{downloads} - $m0->{downloads}, 0, 'nothing crossed PCIe towards the host');
  is($x->at(5, 7), $y->at(5, 7) + $c->at(5, 7), 'host read of a device result');
  my $m2 = ss();
  ok($m2->{downloads} > $m1->{downloads}, 'the read downloaded the buffer (lazy host sync)');
  is(PDL::B200::store_state($x) & 3, 1, 'mirror current and clean after a host READ');
  my $d0 = ss()->{downloads};
  $x->at(9, 9) for 1 .. 5;
  is(ss()->{downloads}, $d0, 'further reads are free');
  my $u0 = ss()->{uploads};
  my $z = $x * 2;
  is(ss()->{uploads}, $u0, 'a device op on a host-READ ndarray uploads nothing');
  $x->set(0, 0, 42);
  is(PDL::B200::store_state($x) & 7, 2, 'host WRITE: mirror modified, device copy stale');
  my $w = $x + 1;
  is(ss()->{uploads}, $u0 + 1, 'the next device op uploads the modified buffer once');
  is($w->at(0, 0), 43, 'and sees the new value');
}
{
  # a 3-op chain is three launches and nothing else
  my ($s0, $m0) = (st(), ss());
  my $r = ($y + $c) * $c - $y;
  my ($s1, $m1) = (st(), ss());
  is($s1->{kernels} - $s0->{kernels}, 3, '3-op chain: 3 launches');
  is($s1->{syncs} - $s0->{syncs}, 0, '3-op chain: no intervening stream sync');
  is($m1->{uploads} + $m1->{downloads} - $m0->{uploads} - $m0->{downloads}, 0, '3-op chain: no PCIe traffic');
  is($s1->{host} - $s0->{host}, 0, '3-op chain: no host fallback');
  is(bytes($r), bytes(cpu(sub { ($y + $c) * $c - $y })), '3-op chain: same bytes as the CPU path');
}
{
  # mixed-type expression: the converttypei the core inserts runs on the device
  my $f = (sequence(float, 3000, 700) % 100) / 4;
  PDL::B200::to_device($f);
  my ($s0, $m0) = (st(), ss());
  my $r = $f + 1.5;
  my ($s1, $m1) = (st(), ss());
  is($r->type . '', 'double', 'float_nd + 1.5 is double');
  is($s1->{host} - $s0->{host}, 0, 'float_nd + 1.5: host_calls == 0');
  is($s1->{adopted} - $s0->{adopted}, 0, 'float_nd + 1.5: nothing migrated');
  is($m1->{downloads} - $m0->{downloads}, 0, 'float_nd + 1.5: no download (the conversion did not run on the CPU)');
  is($s1->{cpu_trans} - $s0->{cpu_trans}, 0, 'float_nd + 1.5: no CPU transformation in between');
  is(bytes($r), bytes(cpu(sub { $f + 1.5 })), 'float_nd + 1.5: same bytes as the CPU path');
  my $l = $f->long;
  is(bytes($l), bytes(cpu(sub { $f->long })), 'explicit ->long on the device');
}
{
  # whole-array reduction of an N-d ndarray: flat = clump(-1) shares the parent's buffer, no copy, no PCIe
  my $x2 = $y + $c;
  my ($s0, $m0) = (st(), ss());
  my $s = $x2->sum;
  my ($s1, $m1) = (st(), ss());
  is($s1->{host} - $s0->{host}, 0, '$x2d->sum: host_calls == 0');
  is($m1->{downloads} - $m0->{downloads}, 0, '$x2d->sum: the ndarray was not downloaded');
  is($m1->{new} + $m1->{recycled} - $m0->{new} - $m0->{recycled}, 0, '$x2d->sum: clump made no copy');
  is($s, cpu(sub { $x2->sum }), '$x2d->sum value');
  my $fl = $x2->flat; $fl->slice('0:9') .= 7;
  is($x2->at(3, 0), 7, 'assignment through flat reaches the parent (two-way dataflow)');
  is($x2->max, cpu(sub { $x2->max }), 'max after the write-through');
}
{
  # slices: in-place ops through a view update the parent's device copy
  my $x = $y + $c;
  my $v = $x->slice('1:-1:2,3:9'); $v += 1000;
  is($x->at(1, 3), $y->at(1, 3) + $c->at(1, 3) + 1000, 'inplace through a slice reaches the parent');
  is($x->at(0, 3), $y->at(0, 3) + $c->at(0, 3), 'elements outside the slice are untouched');
}
{
  # entry points that need plain host memory: the data is handed back to an SV first
  my $x = $y + $c;
  my $ref = $x->get_dataref;
  is(length($$ref), 2048 * 2048 * 8, 'get_dataref of a device result');
  is(unpack('d', substr($$ref, 8 * 5, 8)), $y->at(5, 0) + $c->at(5, 0), 'dataref bytes are current');
  substr($$ref, 0, 8) = pack('d', -1); $x->upd_data;
  is(($x + 0)->at(0, 0), -1, 'upd_data is seen by the next device op');
  my $q = $y + $c; $q->reshape(1024, 4096);
  is($q->at(0, 1), $y->at(1024, 0) + $c->at(1024, 0), 'reshape of a device result');
}
{
  # memory that belongs to someone else is never re-homed: an SV with an outside reference
  my $p = sequence(300, 300) + 0; $p->make_physical;
  PDL::B200::enable(0); my $hostp = sequence(300, 300) * 1; PDL::B200::enable(1);
  my $ref = $hostp->get_dataref;                    # Perl code now holds the SV
  my $t0 = st()->{transient};
  $hostp += 5;
  ok(st()->{transient} > $t0, 'shared SV: staged through a temporary device buffer');
  is(PDL::B200::store_state($hostp), -1, 'shared SV: the ndarray still lives in its SV');
  is(unpack('d', substr($$ref, 8 * 7, 8)), 12, 'shared SV: the result landed in the memory the reference points to');
}
{
  # CPU transformations (not on the device path) see current data through the Core function table
  my $x = ($y + $c)->slice('0:99,0:9');
  my ($s0, $m0) = (st(), ss());
  my $sorted = $x->qsort;                           # PDL::Ufunc::qsort has no device body
  my ($s1, $m1) = (st(), ss());
  ok($s1->{cpu_trans} > $s0->{cpu_trans}, 'qsort was seen as a CPU transformation');
  is($m1->{faults} - $m0->{faults}, 0, 'its input was made current explicitly, not by a fault');
  is(bytes($sorted), bytes(cpu(sub { (($y + $c)->slice('0:99,0:9'))->qsort })), 'qsort of a device result');
  my $str = '' . ($y + $c)->slice('0:2,(0)');
  like($str, qr/^\[1 1\.50\d* 2\.00\d*\]$/, "stringification of a device result: $str");
}
{
  # ->flowing defers the product; sumover as its only consumer runs fused with it (one launch, no 32 MiB intermediate)
  my $a = $y->slice('0:-1:2,(0)')->dummy(1, 1); my $b = $c->slice('0:-1:2,(1)')->dummy(0, 1);
  my $want = ($a * $b)->sumover;
  my ($s0, $m0) = (st(), ss());
  my $got = ($a->flowing * $b)->sumover;
  my ($s1, $m1) = (st(), ss());
  is(PDL::B200::last_kernel(), 'inner', 'fused: the launch is the inner kernel');
  is($m1->{new} + $m1->{recycled} - $m0->{new} - $m0->{recycled}, 0, 'fused: no intermediate buffer');
  is(bytes($got), bytes($want), 'fused: same bytes as the unfused pair');
}
done_testing;
