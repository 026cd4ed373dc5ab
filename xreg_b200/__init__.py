"""xreg_b200 -- B200-native DRR ray casting + 2D similarity metrics behind xReg's
RayCaster / ImgSimMetric2D interfaces.  Thin host layer over libxreg_cuda.so
(include/xreg_cuda.h); there is no CPU or PyTorch fallback."""
from . import _lib
from ._lib import UnsupportedOperationException, XregCudaError, XregError, launch_count
from .geometry import (CameraModel, Volume, downsample_camera_model, exp_se3, se3_inv, to12,
                       kORIGIN_AT_FOCAL_PT_DET_NEG_Z, kORIGIN_AT_FOCAL_PT_DET_POS_Z, kORIGIN_ON_DETECTOR)
from .preproc import downsample_image, downsample_proj_data, log_remap
from .ray_caster import Context, RayCasterDepthCUDA, RayCasterLineIntCUDA, kRAY_CAST_MAX_DEPTH
from .sim_metrics import (ImgSimMetric2D, ImgSimMetric2DCombineMean, ImgSimMetric2DGradNCCCUDA,
                          ImgSimMetric2DNCCCUDA, ImgSimMetric2DPatchGradNCCCUDA, ImgSimMetric2DPatchNCCCUDA,
                          ImgSimMetric2DSSDCUDA, eval_batch)

__all__ = [n for n in dir() if not n.startswith("_")]
