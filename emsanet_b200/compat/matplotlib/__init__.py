"""minimal stand-in: pyplot.get_cmap only (MT/visualization/generic.py:27); see ../README.md"""
__version__ = '0.0-standin'


def use(*args, **kwargs):
    return None
