"""world_size-2 (and 3) CPU tests of the multi-GPU path's host logic over gloo: stream segments +
halo exchange and channel ranges, with the ORACLE standing in for the kernels -- the invariant is
that the partitioned computation equals the un-partitioned one bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle as O
        from solid_dsp_b200 import sharding
        rng = np.random.default_rng(5)  # same stream on every rank; each takes its slice
        if case == "stream":
            T, n, M = 37, 10_007, 4
            h = rng.uniform(-1, 1, T)
            x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
            whole_fir = O.fir_fast(h, x)
            whole_dec = O.fir_fast(h, x, 1.0, M)
            # --- FIR segments (align 1) with a T-1 halo from the previous rank
            first, count = sharding.shard_stream(n, 1, world, rank)
            xl = torch.from_numpy(x[first:first + count].copy())
            halo = torch.zeros(T - 1, dtype=torch.complex64)
            sharding.exchange_halo(xl, halo, rank, world, dist)
            y = O.fir_fast(h, xl.numpy(), hist=halo.numpy())
            assert np.array_equal(y, whole_fir[first:first + count])
            # --- decimator segments: starts are multiples of M -> every rank starts at phase 0
            first, count = sharding.shard_stream(n, M, world, rank)
            assert first % M == 0
            xl = torch.from_numpy(x[first:first + count].copy())
            halo = torch.zeros(T - 1, dtype=torch.complex64)
            sharding.exchange_halo(xl, halo, rank, world, dist)
            y = O.fir_fast(h, xl.numpy(), 1.0, M, count0=0, hist=halo.numpy())
            assert np.array_equal(y, whole_dec[first // M:first // M + len(y)])
            lens = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(lens, torch.tensor([len(y)]))
            assert sum(int(v) for v in lens) == len(whole_dec)  # exact sample count
        elif case == "iir_stream":
            # one long IIR stream in time segments: rank r > 0 warms up over the previous rank's last
            # `warm` samples (outputs discarded) -- within 1e-10 of the unbroken recurrence, no carries
            from solid_dsp_b200.filter.iirdes import stable_lowpass_sections
            ff, fb = stable_lowpass_sections(8)
            n, warm = 20_011, 640  # ||A^640|| < 1e-12 for this cascade (pole radii <= 0.95)
            x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
            whole, _ = O.sos_cascade_fast(ff, fb, x)
            first, count = sharding.shard_stream(n, 32, world, rank)
            assert first % 32 == 0 and count >= warm
            xl = torch.from_numpy(x[first:first + count].copy())
            halo = torch.zeros(warm, dtype=torch.complex64)
            sharding.exchange_halo(xl, halo, rank, world, dist)

            class _OracleIIR:  # the oracle standing in for IIRFilter: reset / execute_block with state
                def __init__(self):
                    self.state = None

                def reset(self):
                    self.state = None

                def execute_block(self, v):
                    y, self.state = O.sos_cascade_fast(ff, fb, np.asarray(v).ravel(), state=self.state)
                    return y

            y = sharding.iir_segment(_OracleIIR(), xl.numpy(), halo.numpy(), rank)
            ref = whole[first:first + count]
            assert np.max(np.abs(y - ref)) <= 1e-9 * np.max(np.abs(ref))
            lens = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(lens, torch.tensor([len(y)]))
            assert sum(int(v) for v in lens) == n
        elif case == "iir_exact":
            # any cascade, also one that does not decay: zero-state segments + one all_gather of the end states +
            # s_{r+1} = A^n s_r + z_r (the oracle stands in for the kernels; A^n from the oracle's own recurrence)
            ff = np.array([0.2, 0.4, 0.2, 1.0, -1.0, 0.0])
            fb = np.array([1.0, -1.9999984, 0.9999984, 1.0, -0.5, 0.25])  # section 0: the reference's active_lag poles (z ~ 1)
            n = 9_001
            x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
            whole, _ = O.sos_cascade_fast(ff, fb, x)
            first, count = sharding.shard_stream(n, 1, world, rank)

            class _OracleIIR:
                def __init__(self):
                    self.state = np.zeros((1, 4), dtype=np.complex128)

                def reset(self):
                    self.state = np.zeros((1, 4), dtype=np.complex128)

                def execute_block(self, v):
                    y, st = O.sos_cascade_fast(ff, fb, np.asarray(v).ravel(), state=self.state[0])
                    self.state = np.asarray(st).reshape(1, 4)
                    return y

                def get_state(self):
                    return self.state.copy(), 0

                def set_state(self, st):
                    self.state = np.asarray(st, dtype=np.complex128).reshape(1, 4)

                def transition(self, k):  # columns = zero-input evolution of the unit states
                    A = np.zeros((4, 4))
                    for j in range(4):
                        e = np.zeros(4, dtype=np.complex128)
                        e[j] = 1.0
                        _, st = O.sos_cascade_fast(ff, fb, np.zeros(k, dtype=np.complex128), state=e)
                        A[:, j] = np.asarray(st).reshape(4).real
                    return A

            y = sharding.iir_segment_exact(_OracleIIR(), torch.from_numpy(x[first:first + count].copy()), rank, world, dist)
            ref = whole[first:first + count]
            assert np.max(np.abs(np.asarray(y) - ref)) <= 1e-5 * np.max(np.abs(whole))  # states cross as complex64
        else:
            Cn, n = 11, 500
            ff = [0.2, 0.4, 0.2, 0.5, 0.0, -0.5]
            fb = [1.0, -0.5, 0.25, 2.0, 0.6, 0.2]
            x = (rng.uniform(-1, 1, (Cn, n)) + 1j * rng.uniform(-1, 1, (Cn, n))).astype(np.complex64)
            first, count = sharding.shard_channels(Cn, world, rank)
            mine = np.stack([O.sos_cascade_fast(ff, fb, x[c])[0] for c in range(first, first + count)]) \
                if count else np.zeros((0, n), dtype=np.complex128)
            spans = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(spans, torch.tensor([first, count]))
            cover = sorted((int(a), int(b)) for a, b in spans)
            pos = 0
            for a, b in cover:  # contiguous, disjoint, complete
                assert a == pos
                pos += b
            assert pos == Cn
            for k, c in enumerate(range(first, first + count)):
                assert np.array_equal(mine[k], O.sos_cascade_fast(ff, fb, x[c])[0])
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", ["stream", "channels", "iir_stream", "iir_exact"])
def test_partitioned_equals_unpartitioned(world, case):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, case, ret), nprocs=world, join=True)
    assert [ret.get(r) for r in range(world)] == ["ok"] * world
