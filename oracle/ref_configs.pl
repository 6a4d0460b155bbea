#!/usr/bin/env perl
# TEST/BENCH INFRASTRUCTURE — times the UNMODIFIED reference (PDL built into oracle/_ref) on bounded versions of
# BASELINE.json configs 1, 3, 4, 5 on the host cores of the box it runs on (SURVEY.md §8(d) "CPU side"); config 2
# is bench.py's cpu_baseline / --impl reference.  Each line: config, size actually timed, ms without and with
# autopthread, and the figure scaled to the full config where the cost model is known (linear, or n^3).
#   perl -Ioracle/_ref/blib/lib -Ioracle/_ref/blib/arch oracle/ref_configs.pl [--quick]
use strict; use warnings;
use PDL::LiteF;
use Time::HiRes qw(time);
use JSON::PP;
my $quick = grep { $_ eq '--quick' } @ARGV;
my $cpus = PDL::Core::online_cpus();
sub timeit { my ($code, $n) = @_; $code->(); my $t0 = time; $code->() for 1 .. $n; return (time - $t0) / $n * 1e3; }
sub both_modes {
  my ($code, $n) = @_;
  PDL::set_autopthread_targ(0); my $t1 = timeit($code, $n);
  PDL::set_autopthread_targ($cpus); PDL::set_autopthread_size(0); my $tn = timeit($code, $n);
  my $act = PDL::get_autopthread_actual();
  PDL::set_autopthread_targ(0);
  return ($t1, $tn, $act);
}
my @out;
{ # cfg1: $x = $y + $c, 2048x2048 double, fresh and preallocated output
  my $y = sequence(2048, 2048) / 1024; my $c = sequence(2048, 2048) * 0.5 + 1; my $x = zeroes(2048, 2048);
  my ($f1, $fn, $a1) = both_modes(sub { my $z = $y + $c; }, 5);
  my ($p1, $pn, $a2) = both_modes(sub { PDL::plus($y, $c, $x, 0); }, 5);
  push @out, { cfg => 'cfg1 $x=$y+$c 2048x2048 double', fresh_ms => $f1, fresh_pthread_ms => $fn, prealloc_ms => $p1,
               prealloc_pthread_ms => $pn, pthreads_actual => $a2, elements_per_sec_best => 2048 * 2048 / (($pn < $p1 ? $pn : $p1) / 1e3) };
}
{ # cfg3: [N,1]*[1,M] on strided slices + dummies, then sumover; timed at N=M=4096 (1/64 of the 32768^2 config), linear in N*M
  my $N = $quick ? 1024 : 4096;
  my ($big1, $big2) = (sequence(2 * $N) / 256, sequence(2 * $N) / 512);
  my $a = $big1->slice('0:-1:2')->dummy(1, 1); my $b = $big2->slice('0:-1:2')->dummy(0, 1);
  my ($t1, $tn, $act) = both_modes(sub { my $s = ($a * $b)->sumover; }, 3);
  my $scale = (32768 / $N) ** 2;
  push @out, { cfg => "cfg3 (a*b)->sumover on slices+dummies, timed at N=M=$N", ms => $t1, pthread_ms => $tn, pthreads_actual => $act,
               full_config_ms_scaled => ($tn < $t1 ? $tn : $t1) * $scale };
}
{ # cfg4: matmult double; a single 2-D matmult never pthreads; n^3 fit from two sizes
  my @ns = $quick ? (256, 512) : (512, 1024);
  my %ms;
  for my $n (@ns) {
    my $A = (sequence($n, $n) % 64 - 32) / 64; my $B = (sequence($n, $n) % 32 - 16) / 16;
    PDL::set_autopthread_targ(0);
    $ms{$n} = timeit(sub { my $C = $A x $B; }, 1);
  }
  my $gf = 2 * $ns[1] ** 3 / ($ms{$ns[1]} / 1e3) / 1e9;
  push @out, { cfg => 'cfg4 matmult double (no pthread possible)', ms => \%ms, gflops => $gf,
               ratio_vs_n3 => ($ms{$ns[1]} / $ms{$ns[0]}) / 8, full_8192_s_scaled => $ms{$ns[1]} / 1e3 * (8192 / $ns[1]) ** 3 };
}
{ # cfg5: sum and max of a 1-D float ndarray (never pthreads: one row); timed at 2^28 (1/32 of 2^33), linear
  my $n = $quick ? 2 ** 22 : 2 ** 28;
  my $x = (sequence(float, $n) % 3) - 1;
  PDL::set_autopthread_targ(0);
  my $ts = timeit(sub { my $s = $x->sum; }, 2); my $tm = timeit(sub { my $m = $x->max; }, 2);
  push @out, { cfg => "cfg5 sum / max of float[2^" . int(log($n) / log(2) + 0.5) . "] 1-D", sum_ms => $ts, max_ms => $tm,
               elements_per_sec_sum => $n / ($ts / 1e3), full_2p33_sum_s_scaled => $ts / 1e3 * (2 ** 33 / $n) };
}
print JSON::PP->new->canonical->encode({ online_cpus => $cpus, reference => "PDL $PDL::VERSION", configs => \@out }), "\n";
