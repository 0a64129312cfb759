"""CPU-only: host addressing of the peer-memory halo transport (hshard.PeerHalo: staging slots by parity / direction,
arrival counters, targets).  The two C-ABI calls are replaced by a ctypes emulation of their documented semantics
(include/dca_b200.h section 5) working on CPU memory, with all pushes of an exchange executed before the waits --
which is what the GPU does in time (a wait spins until the neighbours' pushes have landed)."""
import ctypes

import pytest
import torch

import dcanet_b200 as d

hs = d.hshard
CTAS = 32


class _Emu:
    def __init__(self):
        self.waits = []

    def call(self, name, *a):
        if name == "dca_halo_push":
            t, outer, rows, inner, h, live, up_stage, down_stage, up_flag, down_flag, _ = a
            for stage, flag, first in ((up_stage, up_flag, h), (down_stage, down_flag, rows - h - live)):
                if not stage:
                    continue
                for o in range(outer):
                    ctypes.memmove(stage + o * live * inner, t + (o * rows + first) * inner, live * inner)
                c = ctypes.c_ulonglong.from_address(flag)
                c.value += CTAS
        elif name == "dca_halo_wait_unpack":
            self.waits.append(a)
        elif name == "dca_halo_exchange":                 # one launch: its push half now, its wait half with the others
            self.call("dca_halo_push", *a[:10], a[-1])
            self.waits.append(a[:6] + a[10:])
        else:
            raise AssertionError(name)

    def flush(self):
        for t, outer, rows, inner, h, live, st_top, st_bot, f_top, f_bot, target, err, _ in self.waits:
            for stage, flag, first in ((st_top, f_top, h - live), (st_bot, f_bot, rows - h)):
                if not stage:
                    continue
                assert ctypes.c_ulonglong.from_address(flag).value >= target, "wait would spin forever"
                for o in range(outer):
                    ctypes.memmove(t + (o * rows + first) * inner, stage + o * live * inner, live * inner)
        self.waits = []


def _peers(world, slot):
    bufs = [torch.zeros(4 * slot + 64, dtype=torch.uint8) for _ in range(world)]
    peers = []
    for r in range(world):
        p = hs.PeerHalo.__new__(hs.PeerHalo)
        p.rank, p.world, p.slot, p.buf = r, world, slot, bufs[r]
        p.ptrs = [b.data_ptr() for b in bufs]
        p.ctas, p.epoch = CTAS, 0
        peers.append(p)
    return peers, bufs


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("live", [0, 1])
def test_peer_halo_slots_counters_and_parity(monkeypatch, live, fused):
    world, h = 3, 2
    emu = _Emu()
    monkeypatch.setattr(hs, "FUSED_EXCHANGE", fused)
    monkeypatch.setattr(hs._lib, "call", emu.call)
    monkeypatch.setattr(hs.E, "_stream", lambda: 0)
    peers, bufs = _peers(world, 4096)
    for step in range(3):                                  # three exchanges: both parities, counters keep growing
        ts = [torch.arange(2 * 8 * 4, dtype=torch.float32).view(2, 8, 4) + 1000 * r + 100000 * step
              for r in range(world)]
        want = [t.clone() for t in ts]

        def one(t):
            yield hs.Rows(t, 1, h, fill="keep", live=live)
            return t

        want = hs.drive_lockstep([one(t) for t in want])   # the reference semantics of an exchange
        for r in range(world):
            req = hs.Rows(ts[r], 1, h, fill="keep", live=live)
            assert ts[r].data_ptr() % 16 == 0
            peers[r].refresh(req)
        emu.flush()
        for r in range(world):
            assert torch.equal(ts[r], want[r]), (step, r)
        for r in range(world):                             # arrival counters: one increment of CTAS per exchange
            flags = bufs[r][4 * 4096:4 * 4096 + 16].view(torch.int64)
            assert int(flags[0]) == (CTAS * (step + 1) if r > 0 else 0)
            assert int(flags[1]) == (CTAS * (step + 1) if r < world - 1 else 0)


def test_peer_halo_supports_only_16_byte_rows_that_fit_a_slot():
    peers, _ = _peers(2, 1024)
    p = peers[0]

    class T:                                               # stands in for a CUDA tensor's metadata
        is_cuda = True

        def __init__(self, shape, es=4, ptr=4096):
            self.shape, self._es, self._ptr = shape, es, ptr

        def element_size(self):
            return self._es

        def dim(self):
            return len(self.shape)

        def is_contiguous(self):
            return True

        def data_ptr(self):
            return self._ptr

    assert p.supports(hs.Rows(T((2, 8, 4)), 1, 2))
    assert not p.supports(hs.Rows(T((2, 8, 3)), 1, 2))            # 12-byte rows
    assert not p.supports(hs.Rows(T((2, 8, 4), ptr=4100), 1, 2))  # unaligned base
    assert not p.supports(hs.Rows(T((64, 8, 4)), 1, 2))           # 64 * 2 * 16 B = 2 KB > 1 KB slot
