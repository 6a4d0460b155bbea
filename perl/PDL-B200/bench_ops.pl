#!/usr/bin/env perl
# Times the UNCHANGED operator surface under real PDL with and without the PDL::B200 shim:
#   cfg1  $x = $y + $c on two 2048x2048 double ndarrays (fresh output and preallocated output)
#   chain $z = ($y + $c) * $c - $y   (three ops, intermediates stay on the device)
#   cfg2' sumover/average/minimum of a [16384, 4096] float ndarray with 1% BAD
# Usage: perl -I<pdl blib> -I<shim blib> bench_ops.pl [--reps N]
use strict; use warnings;
use PDL::LiteF;
use PDL::B200 ':noattach';
use Time::HiRes qw(time);
use JSON::PP;
my $reps = 20;
$reps = $ARGV[1] if @ARGV >= 2 && $ARGV[0] eq '--reps';
PDL::set_autopthread_targ(0);

# device ops are asynchronous (no synchronisation between chained ops): the timed region is bracketed by stream
# syncs, so a number is the steady-state time per op INCLUDING the kernels, not the enqueue time
my $gpu = 0;
sub timeit {
  my ($code, $n) = @_;
  $code->() for 1 .. 3;
  PDL::B200::sync() if $gpu;
  my $t0 = time;
  $code->() for 1 .. $n;
  PDL::B200::sync() if $gpu;
  return (time - $t0) / $n * 1e3;
}

my $y = sequence(2048, 2048) / 1024; my $c = sequence(2048, 2048) * 0.5 + 1;
my $x = zeroes(2048, 2048);
my $f = (sequence(float, 16384, 4096) % 17) - 8; $f = $f->setbadif(($f->flat->sequence % 100 == 0)->reshape(16384, 4096));
my %res;
for my $mode (qw(cpu gpu)) {
  if ($mode eq 'gpu') { PDL::B200::attach(); PDL::B200::to_device($_) for ($y, $c, $x, $f); $gpu = 1; }
  $res{$mode}{cfg1_fresh_ms}    = timeit(sub { my $r = $y + $c; 1 }, $reps);
  $res{$mode}{cfg1_prealloc_ms} = timeit(sub { PDL::plus($y, $c, $x, 0); 1 }, $reps);
  $res{$mode}{chain3_ms}        = timeit(sub { my $z = ($y + $c) * $c - $y; 1 }, $reps);
  $res{$mode}{mixed_type_ms}    = timeit(sub { my $z = $f + 1.5; 1 }, $reps > 10 ? 10 : $reps);     # float_nd + 1.5: converttypei + plus
  $res{$mode}{sum_2d_ms}        = timeit(sub { my $z = $y->sum; 1 }, $reps);                          # flat (clump) + sumover + host read of the scalar
  $res{$mode}{sumover_ms}       = timeit(sub { my $s = $f->sumover; 1 }, $reps);
  $res{$mode}{average_ms}       = timeit(sub { my $s = $f->average; 1 }, $reps);
  $res{$mode}{minimum_ms}       = timeit(sub { my $s = $f->minimum; 1 }, $reps);
  $res{$mode}{check} = ($y + $c)->sumover->slice('0:1') . '';
}
$res{stats} = [PDL::B200::stats()];
$res{store_stats} = [PDL::B200::store_stats()];
$res{stats_legend} = 'device calls, host calls, adopted, staged, kernels, binding syncs, transient, cpu transformations';
$res{store_legend} = 'buffers new, recycled, uploads, upload bytes, downloads, download bytes, faults, adopted';
$res{online_cpus} = PDL::Core::online_cpus();
print JSON::PP->new->canonical->encode(\%res), "\n";
