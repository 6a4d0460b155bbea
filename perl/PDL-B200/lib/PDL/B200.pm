package PDL::B200;
# Run PDL's broadcast-loop hot path (PDL::Ops elementwise ops, PDL::Ufunc reductions,
# PDL::Primitive::matmult) on a B200 through libpdlb200, under the UNCHANGED operator surface:
#
#   use PDL::LiteF; use PDL::B200;       # attach() runs at import
#   my $x = $y + $c;                     # pdl_plus_vtable.readdata -> pdlb200_readdata
#   print $x->sumover;                   # host access: pages migrate back on demand
#
# How it attaches (SURVEY.md §8(b) "zero-touch attach"): every pp_def exports its
# `pdl_<op>_vtable`; we look the symbol up in the already-loaded PDL::Ops / PDL::Ufunc /
# PDL::Primitive shared objects and swap the readdata/redodims pointers (B200.xs).
use strict; use warnings;
use PDL::Core ();
use PDL::Ops (); use PDL::Ufunc (); use PDL::Primitive (); use PDL::Bad ();
require DynaLoader;
our @ISA = ('DynaLoader');
our $VERSION = '0.01';
bootstrap PDL::B200 $VERSION;

# op name => PDLB200_OP_* (include/pdlb200.h)
our %OPS = (
  'PDL::Ops' => { plus=>0, mult=>1, minus=>2, divide=>3, gt=>4, lt=>5, le=>6, ge=>7, eq=>8, ne=>9,
    shiftleft=>10, shiftright=>11, or2=>12, and2=>13, xor=>14, power=>15, atan2=>16, modulo=>17,
    spaceship=>18, bitnot=>19, sqrt=>20, sin=>21, cos=>22, not=>23, exp=>24, log=>25, log10=>26,
    _rabs=>27, assgn=>28, abs2=>29, ipow=>62 },
  'PDL::Ufunc' => { sumover=>30, prodover=>31, dsumover=>32, dprodover=>33, average=>34, daverage=>35,
    minimum=>36, maximum=>37, minimum_ind=>38, maximum_ind=>39, andover=>40, orover=>41,
    bandover=>42, borover=>43, zcover=>44, xorover=>45, bxorover=>46,
    cumusumover=>50, cumuprodover=>51, dcumusumover=>52, dcumuprodover=>53,
    minmaximum=>77, magnover=>78 },
  'PDL::Bad' => { nbadover=>47, ngoodover=>48, isbad=>63, isgood=>64, isnan=>65, setbadif=>66, setvaltobad=>67,
    setnantobad=>68, setinftobad=>69, setnonfinitetobad=>70, setbadtonan=>71, setbadtoval=>72, badmask=>73,
    copybad=>74 },
  'PDL::Primitive' => { matmult=>60, axisvalues=>75, inner=>76, outer=>79 },
);

sub _libref {
  my ($module) = @_;
  for my $i (0 .. $#DynaLoader::dl_modules) {
    return $DynaLoader::dl_librefs[$i] if $DynaLoader::dl_modules[$i] eq $module;
  }
  die "PDL::B200: $module is not loaded";
}

our %ATTACHED;
sub attach {
  my %only = map { ($_ => 1) } @_;
  for my $module (sort keys %OPS) {
    my $lib = _libref($module);
    for my $op (sort keys %{ $OPS{$module} }) {
      next if %only && !$only{$op};
      my $addr = DynaLoader::dl_find_symbol($lib, "pdl_${op}_vtable")
        or die "PDL::B200: pdl_${op}_vtable not found in $module";
      _hook($addr, $OPS{$module}{$op});
      $ATTACHED{$op} = 1;
    }
  }
  return scalar keys %ATTACHED;
}

sub import {
  my ($class, @args) = @_;
  return if grep { $_ eq ':noattach' } @args;
  attach() unless %ATTACHED;
}

END { detach() }

1;
