package File::Map;
# Offline stand-in (build of oracle/_ref only): slurp instead of mmap.
use strict; use warnings;
our $VERSION = '0.71';
sub map_file {
  my (undef, $file) = @_;
  open my $fh, '<:raw', $file or die "open $file: $!";
  local $/;
  $_[0] = <$fh>;
  1;
}
1;
