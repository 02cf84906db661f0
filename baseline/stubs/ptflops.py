"""Stand-in for ptflops (models/seg_hrnet.py:38, main.py:162 of the reference): complexity counting is not on the path."""


def get_model_complexity_info(*a, **k):
    return 0, 0
