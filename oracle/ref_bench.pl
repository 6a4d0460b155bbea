#!/usr/bin/env perl
# TEST/BENCH INFRASTRUCTURE — times the UNMODIFIED reference (PDL built into oracle/_ref)
# on a bounded sample of bench.py's workload, on the host cores of the box it runs on.
#
#   perl -Ioracle/_ref/blib/lib -Ioracle/_ref/blib/arch oracle/ref_bench.pl \
#        --n 16384 --rows 2048 --steps 3 --warmup 1 [--threads T]
#
# Workload = BASELINE.json configs[1]: sumover / average / minimum along dim 0 of a
# [n, rows] float ndarray with 1% BAD, values from the same counter hash bench.py uses
# (splitmix64 finaliser of the flat index, seed 0x5EED; SURVEY.md §8(d)).
# --threads 0 => PDL_AUTOPTHREAD_TARG=0 (no pthreads); T>0 => autopthread target T, size 0.
use strict; use warnings;
use PDL::LiteF;
use Time::HiRes qw(time);
use Getopt::Long;
use JSON::PP;

my ($n, $rows, $steps, $warmup, $threads, $row0) = (16384, 2048, 3, 1, 0, 0);
GetOptions('n=i' => \$n, 'rows=i' => \$rows, 'steps=i' => \$steps, 'warmup=i' => \$warmup,
           'threads=i' => \$threads, 'row0=i' => \$row0) or die "bad args";
PDL::set_autopthread_targ($threads);
PDL::set_autopthread_size(0);

sub hash64 {
  my ($z) = @_;                       # ulonglong ndarray; all ops wrap mod 2^64
  $z = $z + pdl(ulonglong, 0x5EED);
  $z = ($z ^ ($z >> 30)) * pdl(ulonglong, '13787848793156543929');   # 0xBF58476D1CE4E5B9
  $z = ($z ^ ($z >> 27)) * pdl(ulonglong, '10723151780598845931');   # 0x94D049BB133111EB
  return $z ^ ($z >> 31);
}

my $a;
{
  my $idx = sequence(ulonglong, $n, $rows) + pdl(ulonglong, $row0) * pdl(ulonglong, $n);
  my $z = hash64($idx);
  my $vals = float((($z >> 11) % 17)) - 8;
  my $isbad = ((($z >> 40) % 100) == 0);
  $a = $vals->setbadif($isbad);
}
$a->make_physical;

my @ops = qw(sumover average minimum);
my %t; my @outs;
for my $it (1 .. $warmup + $steps) {
  for my $op (@ops) {
    my $t0 = time;
    my $o = $a->$op;
    my $dt = time - $t0;
    $t{$op} += $dt if $it > $warmup;
    $outs[0]{$op} = $o if $it == 1;
  }
}
my $total = 0; $total += $t{$_} for @ops;
my $elems = $n * $rows;
my %res = (
  n => $n, rows => $rows, steps => $steps, warmup => $warmup,
  threads_requested => $threads, autopthread_actual => PDL::get_autopthread_actual(),
  online_cpus => PDL::Core::online_cpus(), pdl_version => "$PDL::VERSION",
  ms_per_step => 1000 * $total / $steps,
  elements_per_sec => 3 * $elems * $steps / $total,
  per_op_ms => { map { ($_ => 1000 * $t{$_} / $steps) } @ops },
  checksum => { map { ($_ => $outs[0]{$_}->dsum . '') } @ops },
  nbad_rows => $outs[0]{sumover}->nbad . '',
);
print JSON::PP->new->canonical->encode(\%res), "\n";
