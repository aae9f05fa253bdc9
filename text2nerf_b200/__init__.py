"""text2nerf_b200 -- B200-native TensoRF ray-marching path behind the operator surface of
eckertzhang/Text2NeRF (models.tensoRF.TensorVMSplit / renderer.OctreeRender_trilinear_fast /
dataLoader.ray_utils.get_rays).  See DESIGN.md and INTEGRATION.md."""
from . import _native
from .renderer import OctreeRender_trilinear_fast, SimpleSampler
from .tensorBase import AlphaGridMask, TensorBase, raw2alpha
from .tensoRF import TensorCP, TensorVM, TensorVMSplit

__all__ = ["TensorVMSplit", "TensorBase", "AlphaGridMask", "TensorVM", "TensorCP", "raw2alpha",
           "OctreeRender_trilinear_fast", "SimpleSampler", "_native"]
