"""Host-side mirror of the render driver of the reference's ``renderer.py``:
SimpleSampler (:14-26) and OctreeRender_trilinear_fast (:28-42).  Same signatures and return
tuple; each chunk is one fused-kernel forward of the model.  The chunk loop and the final
concatenation are kept because callers rely on the 5-tuple of dense [N,S] tensors."""
import numpy as np
import torch

from .tensoRF import AlphaGridMask, TensorCP, TensorVM, TensorVMSplit, raw2alpha  # noqa: F401 (re-exported like the reference)


class SimpleSampler:
    """Epoch-wise random permutation of `total` ray indices served in `batch`-sized slices."""

    def __init__(self, total, batch):
        self.total = total
        self.batch = batch
        self.curr = total
        self.ids = None

    def nextids(self):
        self.curr += self.batch
        if self.curr + self.batch > self.total:
            self.ids = torch.LongTensor(np.random.permutation(self.total))
            self.curr = 0
        return self.ids[self.curr:self.curr + self.batch]


def OctreeRender_trilinear_fast(rays, tensorf, chunk=4096, N_samples=-1, ndc_ray=False, white_bg=True,
                                is_train=False, device='cuda'):
    n_rays = rays.shape[0]
    parts = ([], [], [], [])
    if not is_train and white_bg:
        # Evaluation renders draw no random numbers (no jitter, tensorBase.py:313-317; no background draw when white_bg,
        # :497) and rays are independent, so chunk boundaries cannot change any output: callers' small chunks (4096 /
        # 8192 in renderer.evaluation*) are merged into launches that fill the GPU (the kernels reach full throughput
        # from ~65 k rays).  Training calls keep the caller's chunks: each draws torch.rand(R, 1) from the CPU generator.
        chunk = max(int(chunk), 1 << 18)
    for start in range(0, n_rays, chunk):
        rays_chunk = rays[start:start + chunk].to(device, non_blocking=True)
        out = tensorf(rays_chunk, is_train=is_train, white_bg=white_bg, ndc_ray=ndc_ray, N_samples=N_samples)
        for acc, t in zip(parts, out):
            acc.append(t)
    # a single chunk is returned as is: torch.cat would copy the two dense [N,S] tensors (2 x 2.65 GB for an 800x800
    # view at S = 1036) for nothing
    rgbs, depth_maps, z_val, weights = (p[0] if len(p) == 1 else torch.cat(p) for p in parts)
    return rgbs, None, depth_maps, weights, z_val


@torch.no_grad()
def evaluation_views(all_rays, tensorf, img_wh, all_rgbs=None, N_vis=5, N_samples=-1, white_bg=True, ndc_ray=False,
                     push_depth=2.0, device='cuda', chunk=4096):
    """The render loop of renderer.evaluation (renderer.py:81-101, 112-113) without its file / video / LPIPS side: every
    N-th view of `all_rays` [n_views, H*W, 6] is rendered and assembled ON THE DEVICE (clamp, uint8 image, shifted depth,
    PSNR against all_rgbs when given); only the finished uint8 image and the depth map cross to the host, once per view.
    Returns (PSNRs, rgb_maps [list of (H, W, 3) uint8 numpy], depth_maps [list of (H, W) float32 numpy])."""
    from .consumers import assemble_view
    W, H = img_wh
    step = 1 if N_vis < 0 else max(all_rays.shape[0] // N_vis, 1)
    idxs = list(range(0, all_rays.shape[0], step))
    PSNRs, rgb_maps, depth_maps = [], [], []
    for idx in idxs:
        rays = all_rays[idx].view(-1, all_rays.shape[-1])
        rgb_map, _, depth_map, _, _ = OctreeRender_trilinear_fast(rays, tensorf, chunk=chunk, N_samples=N_samples, ndc_ray=ndc_ray,
                                                                 white_bg=white_bg, device=device)
        gt = None if all_rgbs is None else all_rgbs[idx].view(H, W, 3)
        rgb8, depth, psnr = assemble_view(rgb_map, depth_map, H, W, push_depth=push_depth, gt_rgb=gt)
        if psnr is not None:
            PSNRs.append(psnr)
        rgb_maps.append(rgb8.cpu().numpy())
        depth_maps.append(depth.cpu().numpy())
    return PSNRs, rgb_maps, depth_maps
