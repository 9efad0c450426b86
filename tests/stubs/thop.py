"""Test stub of `thop` (only used by the reference's model smoke main)."""
