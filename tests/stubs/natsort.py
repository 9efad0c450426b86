"""Test stub of the `natsort` package (not installed here): natural-order sort, the one function the reference uses."""
import re


def natsorted(seq):
    return sorted(seq, key=lambda s: [int(t) if t.isdigit() else t.lower() for t in re.split(r'(\d+)', str(s))])
