"""`labels`: the fields DS/datasets/cityscapes/cityscapes.py:55-72 indexes (placeholder entries; the Cityscapes dataset
itself is not usable with this stand-in)."""
from collections import namedtuple

Label = namedtuple('Label', 'name id trainId category categoryId hasInstances ignoreInEval color')
labels = [Label(f'c{i}', i, i, 'cat', 0, False, False, (i, i, i)) for i in range(40)]
