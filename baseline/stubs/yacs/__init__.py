"""Stand-in for yacs (config.py:20 of the reference; not installed in this image): the CfgNode surface config.py and
main.py use — attribute access, clone / defrost / freeze, merge_from_file / merge_from_list, dump."""
