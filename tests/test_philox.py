"""Philox4x32-10 known-answer vectors (Random123 kat_vectors) and the derived draws.  CPU only."""
import numpy as np

from oracle import philox as px


def _h(t):
    return [int(x) for x in t]


def test_known_answer_vectors():
    assert _h(px.philox4x32_10(0, 0, 0, 0, 0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert _h(px.philox4x32_10(f, f, f, f, f, f)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _h(px.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_draw_is_keyed_by_global_env_id_only():
    a = px.draw(7, np.arange(0, 64, dtype=np.uint64), 3, 1, px.STREAM_RESET)
    b = px.draw(7, np.arange(32, 64, dtype=np.uint64), 3, 1, px.STREAM_RESET)
    assert np.array_equal(a[32:], b)
    assert not np.array_equal(a, px.draw(8, np.arange(0, 64, dtype=np.uint64), 3, 1, px.STREAM_RESET))


def test_uniform_and_rounded_normal_draws():
    w = px.draw(1, np.arange(200000, dtype=np.uint64), 0, 0, px.STREAM_DEPLOY)[:, 0]
    u = px.u01(w)
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 5e-3
    k3 = px.round_normal(w, 3)
    assert set(np.unique(k3)) == {-3, -2, -1, 0, 1, 2, 3}
    # P(round(N(0,1)) = 0) = Phi(.5) - Phi(-.5) = 0.38292; P(|k| >= 3) = 2 * Phi(-2.5) = 0.01242
    assert abs((k3 == 0).mean() - 0.38292) < 5e-3 and abs((np.abs(k3) == 3).mean() - 0.01242) < 2e-3
    k1 = px.round_normal(w, 1)
    assert set(np.unique(k1)) == {-1, 0, 1} and abs((k1 == 1).mean() - 0.30854) < 5e-3
    assert px.round_normal(np.array([0, 0x0196F4E4, 0x0196F4E5, 0xFFFFFFFF], dtype=np.uint32), 3).tolist() == [-3, -3, -2, 3]


def test_box_muller_is_standard_normal():
    b = px.draw(2, np.arange(200000, dtype=np.uint64), 5, 0, px.STREAM_OBS_NOISE)
    z0, z1 = px.box_muller(b[:, 0], b[:, 1])
    z = np.concatenate([z0, z1])
    assert np.isfinite(z).all() and abs(z.mean()) < 5e-3 and abs(z.std() - 1.0) < 5e-3
