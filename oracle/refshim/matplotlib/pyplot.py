"""Empty stand-in for matplotlib.pyplot."""
