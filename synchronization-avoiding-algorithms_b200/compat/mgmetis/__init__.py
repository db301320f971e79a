"""Stand-in for mgmetis: see compat/README.md."""
