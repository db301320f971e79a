"""Empty stand-in for mgmetis (Data_prepare.py:5); the harness passes epart explicitly."""
