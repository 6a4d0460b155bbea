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
use PDL::Ops (); use PDL::Ufunc (); use PDL::Primitive (); use PDL::Bad (); use PDL::Slices ();
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
    minmaximum=>77, magnover=>78, minimum_n_ind=>90, maximum_n_ind=>91 },
  'PDL::Bad' => { nbadover=>47, ngoodover=>48, isbad=>63, isgood=>64, isnan=>65, setbadif=>66, setvaltobad=>67,
    setnantobad=>68, setinftobad=>69, setnonfinitetobad=>70, setbadtonan=>71, setbadtoval=>72, badmask=>73,
    copybad=>74 },
  'PDL::Primitive' => { matmult=>60, axisvalues=>75, inner=>76, outer=>79 },
);

# flat parent -> child transformations either side of device ops: converttypei (`float_nd + 1.5`) and _clump_int
# (`$x->sum` = flat->sumover); value = [PDLB200_OP_*, kind]
our %FLAT = (
  'PDL::Core'   => { converttypei => [61, 1] },
  'PDL::Slices' => { _clump_int   => [28, 2] },
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
  for my $module (sort keys %FLAT) {
    my $lib = _libref($module);
    for my $op (sort keys %{ $FLAT{$module} }) {
      next if %only && !$only{$op};
      my $addr = DynaLoader::dl_find_symbol($lib, "pdl_${op}_vtable")
        or die "PDL::B200: pdl_${op}_vtable not found in $module";
      _hook($addr, @{ $FLAT{$module}{$op} });
      $ATTACHED{$op} = 1;
    }
  }
  _wrap_perl_side() unless %only;
  return scalar keys %ATTACHED;
}

# Perl-level entry points that refuse or bypass PDL_DONTTOUCHDATA ndarrays: hand the data back to a plain SV
# first (get_dataref: lib/PDL/Core.xs:1145-1160; setdims / reshape: pdlapi.c:183 via pdl_allocdata).
our %WRAPPED;
sub _wrap_perl_side {
  no strict 'refs'; no warnings 'redefine';
  for my $name (qw(get_dataref setdims reshape set_datatype upd_data datasv_refcount)) {
    next if $WRAPPED{$name};
    my $orig = \&{"PDL::$name"};
    $WRAPPED{$name} = $orig;
    *{"PDL::$name"} = sub { PDL::B200::to_host($_[0]) if ref $_[0] && !$_[0]->isnull; goto &$orig; };
  }
  # `($a->flowing * $b)->sumover`: the deferred product and the reduction run as ONE launch (inner)
  if (!$WRAPPED{sumover}) {
    my $orig = \&PDL::sumover;
    $WRAPPED{sumover} = $orig;
    *PDL::sumover = sub {
      if (@_ == 1 && ref $_[0] && (my @p = PDL::B200::_pending_product($_[0]))) {
        $_ = $_->getndims ? $_ : $_->dummy(0) for @p;
        return PDL::inner(@p);
      }
      goto &$orig;
    };
  }
}
sub _unwrap_perl_side {
  no strict 'refs'; no warnings 'redefine';
  *{"PDL::$_"} = delete $WRAPPED{$_} for keys %WRAPPED;
}

sub import {
  my ($class, @args) = @_;
  return if grep { $_ eq ':noattach' } @args;
  attach() unless %ATTACHED;
}

END { _unwrap_perl_side(); detach() }

1;
