package File::Which;
# Offline stand-in (build of oracle/_ref only).
use strict; use warnings;
our $VERSION = '1.27';
require Exporter; our @ISA = ('Exporter');
our @EXPORT = ('which'); our @EXPORT_OK = ('where');
sub which {
  my ($exe) = @_;
  return $exe if $exe =~ m{/} && -x $exe;
  for my $d (split /:/, $ENV{PATH} // '') {
    my $p = "$d/$exe";
    return $p if -x $p && !-d _;
  }
  return;
}
sub where { which(@_) }
1;
