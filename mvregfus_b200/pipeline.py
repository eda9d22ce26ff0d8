"""GPU-resident replacement of the fuse_blockwise block loop (reference: mvregfus/multiview.py:3637-3916) for
the weighted-average fusion with blending weights (fusion_block_overlap = 0).

The reference fuses 128^3 chunks independently and fuse_block makes per-chunk decisions (drop views that are
empty on the chunk, raw copy when fewer than two views remain, drop views whose weights vanish).  Here the whole
volume is fused by ONE kernel launch that applies the same per-chunk decisions, taken from two device probes and
the host probe-point test; no per-block Python loop, no PCIe traffic per block.
"""
import ctypes as C

import numpy as np
import torch

from . import _cabi, fusion
from ._host_workers import probe_any_inside_blocks

CHUNK = 128


def chunk_decisions(data_bits, weight_bits, probe_ok):
    """fuse_block's control flow (mv:3826-3909) per chunk -> chunk_info words.
    data_bits/weight_bits: uint32 per chunk (bit v); probe_ok: bool (nchunks, V)."""
    nv = probe_ok.shape[1]
    info = np.zeros(len(data_bits), dtype=np.uint32)
    for c in range(len(data_bits)):
        live = [v for v in range(nv) if (int(data_bits[c]) >> v) & 1]
        if len(live) == 0:
            info[c] = (1 << 8) | (0 << 12)           # return tviews_block[0]: all zeros
            continue
        if len(live) == 1:
            info[c] = (1 << 8) | (live[0] << 12)     # raw copy of the only live view
            continue
        wl = [v for v in live if probe_ok[c, v] and (int(weight_bits[c]) >> v) & 1]
        if len(wl) == 0:
            info[c] = (1 << 8) | (live[0] << 12)     # tviews_block[0] of the live list
        elif len(wl) == 1:
            info[c] = (1 << 8) | (wl[0] << 12)
        else:
            info[c] = sum(1 << v for v in wl)
    return info


def fuse_volume_blockwise(views, orig_stack_propertiess, params, stack_properties, chunk=CHUNK, out=None):
    """views: list of CUDA tensors (raw views).  Returns the fused uint16 CUDA tensor of stack_properties['size'],
    equal to what fuse_blockwise(..., weights_func=get_weights_simple, fusion_func=fuse_views_weights) produces."""
    lib = _cabi.load()
    dims = [int(s) for s in stack_properties["size"]]
    nch = [-(-d // chunk) for d in dims]
    dev = views[0].device
    vs = fusion.build_viewset(views, orig_stack_propertiess, params, stack_properties, weights="blending", probe=False)
    flags = torch.zeros((2, int(np.prod(nch))), dtype=torch.int32, device=dev)
    st = fusion._stream_ptr()
    for probe in (0, 1):
        _cabi.check(lib.mvf_chunk_probe(vs.n, vs.structs, probe, _cabi.i32x3(dims), _cabi.i32x3((0, 0, 0)),
                                        C.c_void_p(flags[probe].data_ptr()), chunk, _cabi.i32x3(nch), st))
    # probe-point early-out of get_weights_simple on every 128^3 block (host, float64; mv:3491-3532)
    sp = np.asarray(stack_properties["spacing"], dtype=np.float64)
    locs = np.array(list(np.ndindex(*nch)), dtype=np.float64)
    origins = np.asarray(stack_properties["origin"], dtype=np.float64) + locs * chunk * sp
    probe_ok = np.stack([probe_any_inside_blocks(orig_stack_propertiess[v], params[v], origins, [chunk] * 3, sp, 5)
                         for v in range(vs.n)], axis=1)
    bits = flags.cpu().numpy().astype(np.uint32)
    info = torch.from_numpy(chunk_decisions(bits[0], bits[1], probe_ok).astype(np.int32)).to(dev)
    if out is None:
        out = torch.empty(dims, dtype=torch.uint16, device=dev)
    _cabi.check(lib.mvf_fuse_blend_chunked(vs.n, vs.structs, C.c_void_p(out.data_ptr()), _cabi.i32x3(dims),
                                           _cabi.i32x3((0, 0, 0)), C.c_void_p(info.data_ptr()), chunk, _cabi.i32x3(nch), st))
    return out
