package Test::Warn;
# Offline stand-in: just warning_like, as the reference's hot-path tests use it.
use strict; use warnings;
our $VERSION = '0.37';
use Test::Builder;
require Exporter; our @ISA = ('Exporter');
our @EXPORT = qw(warning_like);
my $T = Test::Builder->new;
sub warning_like (&$;$) {
  my ($c, $re, $n) = @_; my @w;
  { local $SIG{__WARN__} = sub { push @w, @_ }; $c->(); }
  my $ok = !defined $re ? !@w
         : ref $re eq 'ARRAY' ? (@w == @$re)
         : (@w == 1 && $w[0] =~ $re);
  $T->ok($ok, $n // 'warning_like') or $T->diag("warnings: @w"); $ok;
}
1;
