"""Host-side mirror of ORB_SLAM_FUSION::OrbExtractor (include/cam/orb_feature/orb_extractor.h:44-104)
over the C ABI.  Same constructor arguments, same outputs (cv::KeyPoint records, N x 32 descriptors,
the monoIndex return value, the public img_pyramid_) -- computed by the sm_100a kernels only."""
import ctypes as C

import numpy as np

from . import _abi as A


def _is_torch(x):
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda")


class OrbExtractor:
    """OrbExtractor(num_feats, scale_factor, num_levs, ini_th_fast, min_th_fast)
    (orb_extractor.cc:407-465).  `device` and `max_batch` are additive, defaulted options."""

    def __init__(self, num_feats, scale_factor, num_levs, ini_th_fast, min_th_fast, device=0, max_batch=1):
        self._lib = A.lib()
        self._h = A.vp()
        self.params = A.Params(int(num_feats), float(scale_factor), int(num_levs), int(ini_th_fast), int(min_th_fast))
        self.device, self.max_batch, self.num_levs = int(device), int(max_batch), int(num_levs)
        rc = self._lib.orbx_create(C.byref(self.params), self.device, self.max_batch, C.byref(self._h))
        if rc:
            self._h = None
            raise A.OrbxError(rc, "orbx_create failed (bad parameters or no CUDA device %d)" % self.device)
        self._shape = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.orbx_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc:
            raise A.OrbxError(rc, self._lib.orbx_last_error(self._h).decode())

    # ---- getters, orb_extractor.h:60-74
    def _table(self, k):
        L = self.num_levs
        arrs = [np.empty(L, np.float32) for _ in range(4)] + [np.empty(L, np.int32)]
        self._check(self._lib.orbx_tables(self._h, *[a.ctypes.data for a in arrs]))
        return arrs[k]

    def GetLevels(self):
        return self.num_levs

    def GetScaleFactor(self):
        return float(np.float32(np.float64(self.params.scale_factor)))

    def GetScaleFactors(self):
        return self._table(0)

    def GetInverseScaleFactors(self):
        return self._table(1)

    def GetScaleSigmaSquares(self):
        return self._table(2)

    def GetInverseScaleSigmaSquares(self):
        return self._table(3)

    def features_per_level(self):
        return self._table(4)

    def max_keypoints(self):
        return self._lib.orbx_max_keypoints(self._h)

    # ---- operator(), orb_extractor.cc:1011-1091
    def __call__(self, img, mask=None, lapping_areas=(0, 0)):
        """Returns (mono_index, keypoints[KP_DTYPE], descriptors[N,32] uint8); mono_index is -1 and
        the outputs are empty for an empty image (:1016).  `mask` is ignored like the reference's."""
        if img is None or getattr(img, "size", 0) == 0:
            return -1, np.empty(0, A.KP_DTYPE), np.empty((0, 32), np.uint8)
        img = np.asarray(img)
        assert img.dtype == np.uint8 and img.ndim == 2, "CV_8UC1 expected (:1019)"
        if img.strides[1] != 1:
            img = np.ascontiguousarray(img)
        h, w = img.shape
        self._shape = (h, w)
        cap = self._lib.orbx_max_keypoints(self._h)
        for _ in range(2):
            kps = np.empty(cap, A.KP_DTYPE)
            desc = np.empty((cap, 32), np.uint8)
            n, nm = C.c_int(0), C.c_int(0)
            rc = self._lib.orbx_extract(self._h, img.ctypes.data, w, h, img.strides[0], int(lapping_areas[0]),
                                        int(lapping_areas[1]), kps.ctypes.data, desc.ctypes.data, cap,
                                        C.byref(n), C.byref(nm))
            if rc == A.E_CAP:
                cap = n.value
                continue
            break
        self._check(rc)    # raises when the second attempt does not fit either
        return nm.value, kps[:n.value].copy(), desc[:n.value].copy()

    def extract_into(self, img, kps, desc, lapping_areas=(0, 0)):
        """operator() into caller-owned arrays (kps [cap] KP_DTYPE, desc [cap,32] uint8): the bare blocking C-ABI call
        orbx_extract, no allocation or copy on the Python side.  Returns (mono_index, n)."""
        h, w = img.shape
        n, nm = C.c_int(0), C.c_int(0)
        self._check(self._lib.orbx_extract(self._h, img.ctypes.data, w, h, img.strides[0], int(lapping_areas[0]),
                                           int(lapping_areas[1]), kps.ctypes.data, desc.ctypes.data, len(kps), C.byref(n), C.byref(nm)))
        return nm.value, n.value

    def extract_begin(self, img, lapping_areas=(0, 0)):
        """First half of operator() (orbx_extract_begin): copies the image and enqueues the pipeline, returns at once.
        Two extractors can so work on the two images of a stereo pair at the same time from one host thread."""
        h, w = img.shape
        self._check(self._lib.orbx_extract_begin(self._h, img.ctypes.data, w, h, img.strides[0], int(lapping_areas[0]),
                                                 int(lapping_areas[1])))

    def extract_end(self, kps, desc):
        """Second half (orbx_extract_end) into caller-owned arrays; returns (mono_index, n)."""
        n, nm = C.c_int(0), C.c_int(0)
        self._check(self._lib.orbx_extract_end(self._h, kps.ctypes.data, desc.ctypes.data, len(kps), C.byref(n), C.byref(nm)))
        return nm.value, n.value

    def extract_batch(self, imgs, lapping_areas=(0, 0), cap=None, stream=None):
        """Batched operator().  `imgs`: [F,H,W] uint8 numpy array (host memory: blocks, returns numpy
        arrays) or CUDA torch tensor (device memory: enqueues on `stream` and returns CUDA tensors).
        Returns (n[F], n_mono[F], kps[F,cap] KP_DTYPE (or [F,cap,7] float32 tensor), desc[F,cap,32])."""
        if _is_torch(imgs):
            import torch
            assert imgs.is_cuda and imgs.dtype == torch.uint8 and imgs.dim() == 3 and imgs.stride(2) == 1
            assert imgs.device.index == self.device, "frames live on cuda:%d, the extractor on cuda:%d" % (imgs.device.index, self.device)
            F, h, w = imgs.shape
            cap = cap or self._cap_for(h, w)
            dev = imgs.device
            kps = torch.empty((F, cap, 7), dtype=torch.float32, device=dev)
            desc = torch.empty((F, cap, 32), dtype=torch.uint8, device=dev)
            n = torch.empty(F, dtype=torch.int32, device=dev)
            nm = torch.empty(F, dtype=torch.int32, device=dev)
            st = stream if stream is not None else A.torch_stream(dev)
            self._check(self._lib.orbx_extract_batch(self._h, imgs.data_ptr(), F, w, h, imgs.stride(1), imgs.stride(0),
                                                     A.MEM_DEVICE, int(lapping_areas[0]), int(lapping_areas[1]),
                                                     kps.data_ptr(), desc.data_ptr(), cap, n.data_ptr(), nm.data_ptr(), st))
            return n, nm, kps, desc
        imgs = np.asarray(imgs)
        assert imgs.dtype == np.uint8 and imgs.ndim == 3 and imgs.strides[2] == 1
        F, h, w = imgs.shape
        explicit_cap = cap is not None   # a caller's own capacity is honoured: frames that do not fit report n[f] = -(needed)
        cap = cap or self._cap_for(h, w)
        for attempt in range(2):
            kps = np.empty((F, cap), A.KP_DTYPE)
            desc = np.empty((F, cap, 32), np.uint8)
            n = np.empty(F, np.int32)
            nm = np.empty(F, np.int32)
            self._check(self._lib.orbx_extract_batch(self._h, imgs.ctypes.data, F, w, h, imgs.strides[1], imgs.strides[0],
                                                     A.MEM_HOST, int(lapping_areas[0]), int(lapping_areas[1]),
                                                     kps.ctypes.data, desc.ctypes.data, cap, n.ctypes.data, nm.ctypes.data, None))
            if F == 0 or n.min() >= 0 or (explicit_cap and n.min() != np.iinfo(np.int32).min):
                break
            # n[f] = -(needed): frame f did not fit `cap`; INT32_MIN: the quadtree's node table overflowed
            if n.min() == np.iinfo(np.int32).min:
                raise A.OrbxError(A.E_UNSUPPORTED, "quadtree node table overflow in frame %d" % int(n.argmin()))
            if attempt == 1:
                raise A.OrbxError(A.E_CAP, "frame %d needs %d keypoint slots, cap is %d" % (int(n.argmin()), -int(n.min()), cap))
            cap = -int(n.min())
        return n, nm, kps, desc

    def extract_batch_into(self, img_ptr, F, w, h, row_stride, frame_stride, mem, lap, kps_ptr, desc_ptr, cap, n_ptr,
                           nm_ptr, stream=None):
        """Thin pass-through of orbx_extract_batch on raw addresses (pinned host buffers, benchmarks)."""
        self._check(self._lib.orbx_extract_batch(self._h, img_ptr, F, w, h, row_stride, frame_stride, mem, int(lap[0]),
                                                 int(lap[1]), kps_ptr, desc_ptr, cap, n_ptr, nm_ptr, stream))

    def _cap_for(self, h, w):
        # orbx_max_keypoints is exact for the geometry of the LAST call and conservative before the first; a new image size
        # can need more than the previous one, so never go below the geometry-free bound (frames that still do not fit
        # report n[f] = -(needed) and extract_batch retries)
        free_bound = sum(max(int(q) + 3, 64) + 2 for q in self.features_per_level())
        return max(self._lib.orbx_max_keypoints(self._h), free_bound) + 8

    def sync(self):
        self._check(self._lib.orbx_sync(self._h))

    def launch_count(self):
        return self._lib.orbx_launch_count(self._h)

    def debug_dropped(self, reset=True):
        """FAST candidates dropped on this device because a list was full (expected 0, see orbx_debug_dropped)."""
        v = C.c_longlong()
        self._check(self._lib.orbx_debug_dropped(self._h, C.byref(v), int(reset)))
        return v.value

    STAGE_NAMES = ("import", "pyramid", "fast_blur", "octree", "describe")

    def set_profiling(self, on=True):
        self._check(self._lib.orbx_set_profiling(self._h, int(on)))

    def stage_times(self, reset=True):
        """({stage: accumulated ms}, chunks) measured with CUDA events on the launching stream."""
        ms = np.zeros(len(self.STAGE_NAMES), np.float64)
        ch = C.c_longlong()
        self._check(self._lib.orbx_stage_times(self._h, ms.ctypes.data, C.byref(ch), int(reset)))
        return dict(zip(self.STAGE_NAMES, ms.tolist())), ch.value

    # ---- ComputePyramid / img_pyramid_, orb_extractor.h:76-78
    def ComputePyramid(self, img):
        img = np.asarray(img)
        assert img.dtype == np.uint8 and img.ndim == 2, "CV_8UC1 expected (:1019)"
        img = np.ascontiguousarray(img)
        h, w = img.shape
        self._check(self._lib.orbx_compute_pyramid(self._h, img.ctypes.data, w, h, img.strides[0]))

    def pyramid_level(self, lev, with_border=False):
        w, h = C.c_int(), C.c_int()
        self._check(self._lib.orbx_pyramid_level(self._h, lev, None, 0, C.byref(w), C.byref(h)))
        buf = np.empty((h.value + 2 * A.EDGE, w.value + 2 * A.EDGE), np.uint8)
        self._check(self._lib.orbx_pyramid_level(self._h, lev, buf.ctypes.data, buf.strides[0], C.byref(w), C.byref(h)))
        return buf if with_border else buf[A.EDGE:-A.EDGE, A.EDGE:-A.EDGE]

    def pyramid_level_size(self, lev):
        w, h = C.c_int(), C.c_int()
        self._check(self._lib.orbx_pyramid_level(self._h, lev, None, 0, C.byref(w), C.byref(h)))
        return w.value, h.value

    @property
    def img_pyramid_(self):
        """Level images of the last single-frame call; like the reference's Mats they are views into
        buffers that carry the 19-px REFLECT_101 border (orb_extractor.cc:1109-1114)."""
        return [self.pyramid_level(l) for l in range(self.num_levs)]

    # ---- stage intermediates (parity tests)
    def stage(self, stage, lev, frame=0):
        cnt = C.c_int()
        self._check(self._lib.orbx_stage_download(self._h, frame, stage, lev, None, 0, C.byref(cnt)))
        if stage in (A.STAGE_LEVEL, A.STAGE_BLUR):
            w, h = C.c_int(), C.c_int()
            self._check(self._lib.orbx_pyramid_level(self._h, lev, None, 0, C.byref(w), C.byref(h)))
            out = np.empty((h.value, w.value), np.uint8)
        else:
            out = np.empty((max(cnt.value, 1), 3), np.int32)
        self._check(self._lib.orbx_stage_download(self._h, frame, stage, lev, out.ctypes.data, out.nbytes, C.byref(cnt)))
        return out if stage in (A.STAGE_LEVEL, A.STAGE_BLUR) else out[:cnt.value]


def synth_frames(kind, n_frames, w, h, seed=1, first_frame=0, shift_x=0, noise_seed=None, device=0, out=None):
    """Deterministic synthetic frames generated on the GPU: kind 'blocks' or 'uniform'.
    Returns a [F,H,W] uint8 CUDA tensor."""
    import torch
    dev = torch.device("cuda", device)
    if out is None:
        out = torch.empty((n_frames, h, w), dtype=torch.uint8, device=dev)
    rc = A.lib().orbx_synth_frames(device, 0 if kind == "blocks" else 1, out.data_ptr(), n_frames, w, h, out.stride(1),
                                   out.stride(0), seed, first_frame, shift_x, seed if noise_seed is None else noise_seed,
                                   A.torch_stream(dev))
    if rc:
        raise A.OrbxError(rc, "orbx_synth_frames")
    return out
