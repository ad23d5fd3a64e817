"""Shared helpers for the test-suite (test infrastructure, not product code)."""
