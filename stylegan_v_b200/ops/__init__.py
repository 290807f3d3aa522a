"""`torch_utils.ops`-compatible operator package backed by libsgv_b200 (sm_100a)."""
from . import upfirdn2d, bias_act, conv2d_gradfix, conv2d_resample, fma, grid_sample_gradfix  # noqa: F401
