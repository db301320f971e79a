"""Stand-in for matplotlib (used only when the real package is not installed): plotting calls are accepted and
ignored, `savefig` writes nothing.  See compat/README.md."""
