"""ORACLE — test infrastructure only. See oracle/pointops_ref.c header."""
