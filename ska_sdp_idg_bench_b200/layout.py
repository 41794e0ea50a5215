"""Array layouts of the reference (app/common/types.hpp, parameters.hpp) as numpy dtypes."""
import numpy as np

NR_CORRELATIONS = 4  # app/common/parameters.hpp:3
IMAGE_SIZE = 0.01  # app/common/parameters.hpp:4
W_STEP = 0.0  # app/common/parameters.hpp:5

# idg::Metadata, app/common/types.hpp:19-26 (36 bytes)
METADATA_DTYPE = np.dtype(
    [
        ("baseline_offset", "<i4"),
        ("time_offset", "<i4"),
        ("nr_timesteps", "<i4"),
        ("aterm_index", "<i4"),
        ("station1", "<u4"),
        ("station2", "<u4"),
        ("x", "<i4"),
        ("y", "<i4"),
        ("z", "<i4"),
    ]
)
assert METADATA_DTYPE.itemsize == 36

# idg::Baseline, app/common/types.hpp:15-17
BASELINE_DTYPE = np.dtype([("station1", "<u4"), ("station2", "<u4")])
