"""Oracles (test infrastructure only): CPU restatements of the reference path used by tests/, __graft_entry__.smoke() and
bench.py's CPU legs as the checker.  The product never imports this package."""
