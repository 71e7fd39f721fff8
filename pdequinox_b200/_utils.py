"""Host-side helpers mirrored from pdequinox/_utils.py (count_parameters :58-63, sum_receptive_fields :240-267)."""
from ._module import count_parameters  # noqa: F401


def sum_receptive_fields(receptive_fields):
    num_spatial_dims = len(receptive_fields[0])
    return tuple(
        (sum(r[i][0] for r in receptive_fields), sum(r[i][1] for r in receptive_fields))
        for i in range(num_spatial_dims))
