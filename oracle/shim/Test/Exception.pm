package Test::Exception;
# Offline stand-in: just the calls the reference's hot-path tests make.
use strict; use warnings;
our $VERSION = '0.43';
use Test::Builder;
require Exporter; our @ISA = ('Exporter');
our @EXPORT = qw(throws_ok lives_ok dies_ok lives_and);
my $T = Test::Builder->new;
sub throws_ok (&$;$) {
  my ($c, $re, $n) = @_;
  eval { $c->() }; my $e = $@;
  my $ok = ref $re eq 'Regexp' ? ($e =~ $re) : (ref $e && $e->isa($re));
  $T->ok($ok, $n // 'threw') or $T->diag("got: $e");
  $@ = $e; $ok;
}
sub dies_ok (&;$) { my ($c, $n) = @_; eval { $c->() }; $T->ok(!!$@, $n // 'died') }
sub lives_ok (&;$) {
  my ($c, $n) = @_; eval { $c->() }; my $e = $@;
  $T->ok(!$e, $n // 'lived') or $T->diag("died: $e"); !$e;
}
sub lives_and (&;$) {
  my ($c, $n) = @_; eval { $c->() };
  if ($@) { $T->ok(0, $n); $T->diag("died: $@") }
}
1;
