"""CPU check of the float32 identities K1b's decision stage relies on (ysmr_b200/csrc/frontend.cu, gd_decide):

    R = mean + 1.5 * 2^23                 -> integer-valued float: rint(mean), ties to even (cv2's saturate_cast after the filter)
    e = b - R                             -> exact (both integers below 2^24)
    d = sat(e + (1.5 * 2^23 - t))         -> 1.0 iff  b - rint(mean) > t   (cv2.adaptiveThreshold's compare)
    byte = low mantissa byte of 2^23 + sum d_k 2^k

numpy float32 arithmetic rounds exactly like the device's add.rn.f32 / fma.rn.f32, so the identity can be verified
exhaustively over b and t on adversarial means without a GPU."""
import numpy as np

MAGIC = np.float32(12582912.0)          # 1.5 * 2^23


def _decide(mean, b, t):
    mean = mean.astype(np.float32); b = b.astype(np.float32)
    r = mean + MAGIC
    e = b - r
    c = MAGIC - np.float32(t)
    return np.clip(e + c, np.float32(0), np.float32(1))


def _adversarial_means():
    ks = np.arange(0, 256, dtype=np.float32)
    halves = ks + np.float32(0.5)
    vals = [ks, halves, np.nextafter(halves, np.float32(0)), np.nextafter(halves, np.float32(1000)),
            np.nextafter(ks, np.float32(0)), np.nextafter(ks, np.float32(1000))]
    rng = np.random.default_rng(0)
    vals.append(rng.uniform(0, 255.5, 200000).astype(np.float32))
    m = np.concatenate(vals)
    return m[(m >= 0) & (m < 255.5)]


def test_decision_equals_integer_compare_exhaustively():
    m = _adversarial_means()
    ref_mean = np.rint(m).astype(np.int32)             # half to even, like cvRound / saturate_cast<uchar>
    for t in (-255, -8, -7, -5, -3, 0, 3, 5, 7, 8, 254):
        for b in range(256):
            got = _decide(m, np.full(m.shape, b), t)
            want = (b - ref_mean) > t
            assert set(np.unique(got)) <= {0.0, 1.0}
            assert (got.astype(bool) == want).all(), (t, b)


def test_difference_and_sum_are_exact():
    m = _adversarial_means()
    r = m + MAGIC
    assert (r == np.rint(m) + np.float64(MAGIC)).all()          # the float32 add IS the rounding to an integer
    for b in (0, 1, 127, 255):
        e = np.float32(b) - r
        assert (e.astype(np.float64) == b - r.astype(np.float64)).all()


def test_byte_packing_in_the_mantissa():
    rng = np.random.default_rng(1)
    d = rng.integers(0, 2, (10000, 8)).astype(np.float32)        # mask bits of px0..3, marker bits of px0..3
    # kernel order: even pixels in .x, odd pixels in .y, broadcast weights 1, 4, 16, 64; byte = x + 2 y on top of 2^23
    x = np.float32(8388608.0) + d[:, 0]
    y = d[:, 1].copy()
    x = d[:, 2] * np.float32(4) + x; y = d[:, 3] * np.float32(4) + y
    x = d[:, 4] * np.float32(16) + x; y = d[:, 5] * np.float32(16) + y
    x = d[:, 6] * np.float32(64) + x; y = d[:, 7] * np.float32(64) + y
    acc = (y * np.float32(2) + x).astype(np.float32)
    byte = acc.view(np.uint32) & 0xFF
    want = sum(d[:, k].astype(np.uint32) << k for k in range(8))
    assert (byte == want).all()


def test_nibble_squeeze_of_the_pack_kernel():
    def squeeze(v):
        v = (v | (v >> 4)) & 0x00FF00FF
        return (v | (v >> 8)) & 0x0000FFFF
    rng = np.random.default_rng(2)
    by = rng.integers(0, 256, (5000, 8)).astype(np.uint64)
    lo = sum(by[:, i] << (8 * i) for i in range(4)); hi = sum(by[:, 4 + i] << (8 * i) for i in range(4))
    mask = squeeze(lo & 0x0F0F0F0F) | (squeeze(hi & 0x0F0F0F0F) << 16)
    mark = squeeze((lo >> 4) & 0x0F0F0F0F) | (squeeze((hi >> 4) & 0x0F0F0F0F) << 16)
    want_mask = sum((by[:, i] & 0xF) << (4 * i) for i in range(8))
    want_mark = sum((by[:, i] >> 4) << (4 * i) for i in range(8))
    assert (mask == want_mask).all() and (mark == want_mark).all()
