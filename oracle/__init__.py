"""Checkers only (test infrastructure). Nothing in apd_mvs_b200/ may import this package."""
