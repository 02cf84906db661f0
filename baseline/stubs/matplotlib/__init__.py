"""Stand-in for matplotlib (main.py:23 of the reference imports pyplot and never plots on the training path)."""
