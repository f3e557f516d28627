"""CPU: building blocks of the window-chain oracle (oracle/window_oracle.py) that do not need a device: the Philox4x32-10
generator against the Random123 known-answer vectors, the 24-bit uniforms, the float32 class mapping, the grid offsets (twin of
mpp_window_grid) and the inverse-CDF consistency check."""
import numpy as np

from oracle import window_oracle as wo


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors: philox4x32 10 rounds, counter then key
    assert wo.philox4x32_10(0, 0, 0, 0, 0) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert wo.philox4x32_10((0xffffffff << 32) | 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff) == \
        (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    # key = (0xa4093822, 0x299f31d0): the 64-bit seed holds key[0] in its low word
    assert wo.philox4x32_10((0x299f31d0 << 32) | 0xa4093822, 0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_uniforms_and_classes():
    assert 0.0 < float(wo.u01f(0)) < 1e-7 and 1.0 - 1e-7 < float(wo.u01f(0xffffffff)) <= 1.0  # (2^24 - 1 + 0.5 rounds up in float32)
    assert wo.u01f(0x80000000).dtype == np.float32
    for i, vmax in enumerate((32.0, 1.0, np.pi)):
        assert wo.class_f32(i, 0.0) == 0 and wo.class_f32(i, np.float32(vmax)) == 31
        for c in (1, 7, 31):
            edge = wo._EDGES32[i][c]
            assert wo.class_f32(i, edge) == c and wo.class_f32(i, np.nextafter(edge, np.float32(-1))) == c - 1


def test_grid_offset_matches_the_host_twin():
    from mpp_cnn_rs_object_detection_b200 import multi_gpu as mg
    for seed, sweep in ((0, 0), (3, 17), (2 ** 62 - 1, 2 ** 33 + 5)):
        assert wo.grid_offset(seed, sweep) == mg.grid_offset(seed, sweep)
        ox, oy = wo.grid_offset(seed, sweep)
        assert 0 <= ox < 32 and 0 <= oy < 32


def test_cdf_consistency_check():
    w = np.array([0.0, 2.0, 0.0, 1.0, 1.0])
    ok = wo.WindowOracle._cdf_consistent
    assert ok(w, 1, 0.25) and ok(w, 1, 0.4999) and not ok(w, 1, 0.6)
    assert ok(w, 3, 0.6) and ok(w, 4, 0.9) and not ok(w, 0, 0.0) and not ok(w, 2, 0.5)
