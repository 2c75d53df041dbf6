"""Multi-GPU partitioning of the filtering path.  One process per GPU.

  * independent channels (decimator / interpolator / batched IIR): contiguous channel ranges, no
    data-path collective at all;
  * one long FIR stream: contiguous time segments; rank r > 0 needs the T-1 samples that precede
    its segment (the "halo") -- one tiny point-to-point message per boundary per step.  For a
    decimator the segment starts are multiples of M so every rank starts at decimator phase 0.
  * one long IIR stream (second-order cascade whose memory decays): contiguous time segments; rank
    r > 0 warms its filter up over the `decay_length` samples that precede its segment (same halo
    exchange, the warm-up output is discarded) -- no carry exchange, within 1e-10 of the unbroken
    recurrence;
  * one long IIR stream, ANY second-order cascade (also one that does not decay, e.g. the reference's
    pll::active_lag loop filter with its double pole at z ~ 1): iir_segment_exact -- every rank runs its segment
    from zero state, the ranks all-gather their end states (2 * sections complex values each: the only
    collective on the whole path), every rank evaluates s_{r+1} = A^{n_r} s_r + z_r in f64
    with A^n from sgpu_iir_transition, and ranks r > 0 re-run their segment from the true start state.

The arithmetic lives in the C ABI (sgpu_shard_channels / sgpu_shard_stream); the exchange uses
torch.distributed point-to-point ops so the same code runs over NCCL (GPU) and gloo (CPU tests)."""
from __future__ import annotations

import ctypes as C

from . import _ffi


def shard_channels(n_channels: int, world: int, rank: int):
    first, count = _ffi.c_size(), _ffi.c_size()
    _ffi.check(_ffi.lib.sgpu_shard_channels(n_channels, world, rank, C.byref(first), C.byref(count)))
    return first.value, count.value


def shard_stream(n_samples: int, align: int, world: int, rank: int):
    first, count = _ffi.c_size(), _ffi.c_size()
    _ffi.check(_ffi.lib.sgpu_shard_stream(n_samples, align, world, rank, C.byref(first), C.byref(count)))
    return first.value, count.value


def exchange_halo(x_local, halo_out, rank: int, world: int, dist):
    """Send the last len(halo_out) samples of this rank's segment to rank+1 and receive the
    previous rank's tail into halo_out (rank 0 keeps halo_out as is: the filter's own history).
    x_local / halo_out are 1-D complex64 torch tensors on the communicator's device."""
    h = halo_out.shape[0]
    if world == 1 or h == 0:  # a 1-tap filter has no history: nothing to exchange (x_local[-0:] would be the whole segment)
        return halo_out
    if rank + 1 < world and x_local.shape[0] < h:
        raise ValueError(f"rank {rank}: segment of {x_local.shape[0]} samples is shorter than the halo ({h}): the next "
                         "rank would need samples of the rank before this one; use fewer ranks for this stream")
    ops = []
    if rank + 1 < world:
        tail = x_local[-h:].contiguous()
        ops.append(dist.P2POp(dist.isend, tail, rank + 1))
    if rank > 0:
        ops.append(dist.P2POp(dist.irecv, halo_out, rank - 1))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return halo_out


def iir_segment(filt, x_local, halo, rank: int):
    """Run this rank's segment of one long IIR stream: rank r > 0 first resets the filter and runs
    it over the halo (the previous rank's last decay_length samples), discarding that output."""
    if rank > 0:
        filt.reset()
        filt.execute_block(halo)
    return filt.execute_block(x_local)


def iir_segment_exact(filt, x_local, rank: int, world: int, dist):
    """Exact time-segmented second-order cascade for any filter.

    filt: IIRFilter (SecondOrder, one channel) on every rank; rank 0's handle holds the stream's entry state.
    x_local: this rank's contiguous segment.  Returns this rank's outputs; afterwards the LAST rank's handle holds
    the stream's end state.  One collective: an all_gather of 2 * sections complex values (+ the segment length)."""
    import numpy as np
    import torch
    if world == 1:
        return filt.execute_block(x_local)
    if rank > 0:
        filt.reset()
    y = filt.execute_block(x_local)  # rank 0: the true outputs; rank r > 0: zero-state run, only its end state is kept
    z, _ = filt.get_state()
    D = z.shape[-1]
    n_loc = int(x_local.shape[-1])
    dev = x_local.device if hasattr(x_local, "device") and not isinstance(x_local, np.ndarray) else torch.device("cpu")
    mine = torch.zeros(2 * D + 1, dtype=torch.float64, device=dev)
    mine[:D] = torch.from_numpy(z[0].real.astype(np.float64))
    mine[D:2 * D] = torch.from_numpy(z[0].imag.astype(np.float64))
    mine[2 * D] = float(n_loc)
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    if rank == 0:
        return y
    allv = [v.cpu().numpy() for v in allv]
    s = allv[0][:D] + 1j * allv[0][D:2 * D]  # end state of rank 0 = start state of rank 1 (exact: run from the true state)
    for r in range(1, rank):
        A = filt.transition(int(allv[r][2 * D]))
        s = A @ s + (allv[r][:D] + 1j * allv[r][D:2 * D])
    filt.set_state(s[None, :])  # the handle rounds to its own state precision (f32 on the device)
    return filt.execute_block(x_local)
