"""Empty stand-in for matplotlib (DNN_tools.py:7 imports pyplot at module top)."""
