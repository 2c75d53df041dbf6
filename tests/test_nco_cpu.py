"""CPU tests of the NCO restatement (oracle/solid_oracle.c, nco/mod.rs).  The reference holds no golden for the NCO,
so the oracle is checked against an independent pure-Python transcription of the same source lines (small cases) and
against closed forms of the 32-bit phase arithmetic; the product's host-side `constrain` must agree with both."""
import math

import numpy as np

import oracle as O

TWO_PI = 2.0 * math.pi


class PyNCO:
    """Line-by-line Python transcription of nco/mod.rs:26-172 (u32 arithmetic made explicit)."""

    def __init__(self):
        self.table = [math.sin(TWO_PI * i / 1024.0) for i in range(1024)]  # :37-40
        self.theta = 0
        self.delta = 0

    @staticmethod
    def constrain(theta):  # :176-188
        frac = math.modf(theta / TWO_PI)[0]
        if frac < 0.0:
            frac += 1.0
        v = frac * float(0xFFFFFFFF)
        return max(0, min(0xFFFFFFFF, int(v)))  # `as u32`: truncation, saturating

    def step(self):  # :93-96
        self.theta = (self.theta + self.delta) & 0xFFFFFFFF

    def index(self):  # :98-101
        return (((self.theta + (1 << 21)) & 0xFFFFFFFF) >> 22) & 0x3FF

    def sin(self):
        return self.table[self.index()]

    def cos(self):
        return self.table[(self.index() + 256) & 0x3FF]

    def mix_down(self, x):  # :147-151
        return complex(self.cos(), self.sin()).conjugate() * x

    def mix_up(self, x):  # :141-145
        return complex(self.cos(), self.sin()) * x


def test_constrain_matches_transcription_and_product():
    from solid_dsp_b200.nco import constrain  # host arithmetic of the product library, no GPU involved
    cases = [0.0, 0.1, -0.1, 1.0, math.pi, -math.pi, TWO_PI, -TWO_PI, 7.0, 100.5, -1234.5678, 1e-12, -1e-12, 6.283185307179586 - 1e-9]
    for t in cases:
        assert O.nco_constrain(t) == PyNCO.constrain(t) == constrain(t), t
    assert O.nco_constrain(0.1) == 68356527  # main.rs:30 `nco.set_frequency(0.1)`: 0.1 / 2 pi of a turn, truncated


def test_sincos_sequence_of_main_rs():
    """main.rs:29-36: set_frequency(0.1), then sincos() / step() in a loop."""
    n, p = O.NCO(), PyNCO()
    n.set_frequency(0.1)
    p.delta = PyNCO.constrain(0.1)
    d = p.delta
    for i in range(5000):
        assert n.raw() == (p.theta, p.delta)
        assert n.sincos() == (p.sin(), p.cos())
        # closed form of the accumulator and of the rounded table index
        assert p.theta == (i * d) & 0xFFFFFFFF
        k = ((p.theta + (1 << 21)) & 0xFFFFFFFF) >> 22
        assert n.sin() == math.sin(TWO_PI * k / 1024.0)
        n.step()
        p.step()


def test_phase_and_frequency_adjust_wrap():
    n = O.NCO()
    n.set_phase(-0.25)
    n.set_frequency(6.0)
    th, dl = n.raw()
    assert th == PyNCO.constrain(-0.25) and dl == PyNCO.constrain(6.0)
    n.adjust_frequency(5.0)   # 6.0 + 5.0 > 2 pi: the u32 sum wraps
    n.adjust_phase(6.2)
    assert n.raw() == ((th + PyNCO.constrain(6.2)) & 0xFFFFFFFF, (dl + PyNCO.constrain(5.0)) & 0xFFFFFFFF)
    n.reset()
    assert n.raw() == (0, 0)


def test_mix_block_is_mix_then_step():
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, 300) + 1j * rng.uniform(-1, 1, 300)
    for up in (False, True):
        n, p = O.NCO(), PyNCO()
        n.set_raw(0xFFF00000, 0x01234567)   # wraps within the block
        p.theta, p.delta = 0xFFF00000, 0x01234567
        got = n.mix_up_block(x) if up else n.mix_down_block(x)
        ref = []
        for v in x:
            ref.append(p.mix_up(v) if up else p.mix_down(v))
            p.step()
        assert np.array_equal(got, np.array(ref))
        assert n.raw() == (p.theta, p.delta)
    # the helper used by bench.py and the GPU tests: rows are independent NCOs
    y = O.nco_mix_down_block(np.stack([x, x]), raw=[(1, 2), (0xFFF00000, 0x01234567)])
    q = PyNCO()
    q.theta, q.delta = 0xFFF00000, 0x01234567
    ref = []
    for v in x:
        ref.append(q.mix_down(v))
        q.step()
    assert np.array_equal(y[1], np.array(ref))


def test_ddc_fast_is_mix_then_decimate():
    rng = np.random.default_rng(6)
    h = rng.uniform(-1, 1, 23)
    x = rng.uniform(-1, 1, 400) + 1j * rng.uniform(-1, 1, 400)
    n = O.NCO()
    n.set_frequency(0.37)
    d = O.DecimatingFIRFilter(h, 0.5, 4)
    ref = []
    for v in x:  # the loop a user of the reference writes
        ref.extend(d.execute(n.mix_down(v)))
        n.step()
    assert np.array_equal(O.ddc_fast(h, x, 0.5, 4, 0.37), np.array(ref))
