"""Primitive table and the shipped cell genotype (reference: modeling/genotypes.py:5-14 and
searched_arch/autodeeplab/genotype.npy — the file every reference script loads, eval.py:43,68)."""
import collections

import numpy as np

# index -> op name, in the order the genotype's primitive indices refer to (reference table, genotypes.py:5-14)
PRIMITIVES = ['none'] + [f'{kind}_pool_3x3' for kind in ('max', 'avg')] + ['skip_connect'] + \
             [f'{kind}_conv_{k}x{k}' for kind in ('sep', 'dil') for k in (3, 5)]
assert PRIMITIVES[4:] == ['sep_conv_3x3', 'sep_conv_5x5', 'dil_conv_3x3', 'dil_conv_5x5']

Genotype = collections.namedtuple('Genotype', ['cell', 'cell_concat'])

# rows = [branch_index, primitive_index]; int64 [10, 2] exactly as stored in genotype.npy
AUTODEEPLAB_CELL = np.array(
    [[0, 7], [1, 4], [2, 4], [3, 6], [5, 4], [8, 4], [11, 5], [13, 5], [19, 7], [18, 5]], dtype=np.int64)

# network paths hard-coded in the reference drivers (eval.py:42-84): name -> C -> (network_arch, C_index, low_level_layer)
NETWORKS = {
    'searched-dense': {
        2: ([1, 2, 2, 2, 3, 2, 2, 1, 1, 1, 1, 2], [5], 0),
        3: ([1, 2, 3, 2, 2, 3, 2, 3, 2, 3, 2, 3], [3, 7], 0),
        4: ([1, 2, 3, 3, 2, 3, 3, 3, 3, 3, 2, 2], [2, 5, 8], 0),
    },
    'autodeeplab-dense': {
        2: ([0, 0, 0, 1, 2, 1, 2, 2, 3, 3, 2, 1], [5], 2),
        3: ([0, 0, 0, 1, 2, 1, 2, 2, 3, 3, 2, 1], [3, 7], 2),
        4: ([0, 0, 0, 1, 2, 1, 2, 2, 3, 3, 2, 1], [2, 5, 8], 2),
    },
}
