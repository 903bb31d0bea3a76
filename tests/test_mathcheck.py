"""The product's __host__ __device__ arithmetic (if-defense_b200/csrc/{ifd_math,convonet_point}.cuh),
instantiated on the host by tests/mathcheck and driven like the kernels, against the golden fixtures.
This catches formula errors on CPU; the -m gpu tests then check the parallel glue on the device."""
import ctypes

import numpy as np

D, F32 = ctypes.c_double, ctypes.c_float


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def planes_cl(conv):
    return np.ascontiguousarray(conv["planes_nchw"].transpose(0, 1, 3, 4, 2))


def test_decode_fwd_bwd(mathcheck, conv):
    pl, blob = planes_cl(conv), conv["dec_blob"]
    for pk, lk, gk in (("p0", "logits", "grad_p"), ("clamp_p", "clamp_logits", "clamp_grad_p")):
        xyz = np.ascontiguousarray(conv[pk])
        B, K, _ = xyz.shape
        logits, grad = np.zeros((B, K), np.float32), np.zeros((B, K, 3), np.float32)
        mathcheck.mc_convonet_decode(P(blob), P(pl), P(xyz), B, K, 64, 5, D(0.1), 1, P(conv["gl"]), F32(0), F32(0), P(logits), P(grad))
        assert np.abs(logits - conv[lk]).max() < 2e-6
        scale = np.abs(conv[gk]).max()
        assert np.abs(grad - conv[gk]).max() < 2e-6 * scale


def test_knn_repulsion(mathcheck, geo):
    x = np.ascontiguousarray(geo["xyz"])
    B, K, _ = x.shape
    idx, loss, grad = np.zeros((B, K, 5), np.int32), np.zeros(B, np.float32), np.zeros((B, K, 3), np.float32)
    mathcheck.mc_knn_repulsion(P(x), B, K, 5, F32(0.07), F32(0.03), F32(1e-12), P(geo["rep_grad_loss"]), P(idx), P(loss), P(grad), None)
    assert np.array_equal(idx, geo["knn5"])                                 # bit-exact indices
    np.testing.assert_allclose(loss, geo["rep_loss"], rtol=2e-6)
    assert np.abs(grad - geo["rep_grad"]).max() < 2e-6 * np.abs(geo["rep_grad"]).max()


def test_duplicate_points(mathcheck, geo):
    x = np.ascontiguousarray(geo["dup_xyz"])
    loss, grad = np.zeros(1, np.float32), np.zeros((1, 64, 3), np.float32)
    mathcheck.mc_knn_repulsion(P(x), 1, 64, 5, F32(0.07), F32(0.03), F32(1e-12), None, None, P(loss), P(grad), None)
    np.testing.assert_allclose(loss, geo["dup_rep_loss"], rtol=2e-6)
    assert np.abs(grad - geo["dup_rep_grad"]).max() < 2e-6 * np.abs(geo["dup_rep_grad"]).max()


def test_long_accumulator_is_exact_and_order_free(mathcheck, geo):
    """fx_term/fx_value: tiny and huge terms survive together and the order does not matter."""
    import subprocess, os, tempfile, textwrap
    src = textwrap.dedent('''
        #include <stdio.h>
        #include <stdlib.h>
        #include "%s"
        using namespace ifd;
        int main() {
          float t[6] = {1.5f, -1.5f, 3e-30f, 1e-38f, -2e-30f, 7.25f};
          long long a[kFxLimbs] = {0}, b[kFxLimbs] = {0};
          for (int i = 0; i < 6; ++i) { FxTerm x = fx_term(t[i]); if (x.limb >= 0) a[x.limb] += x.val; }
          for (int i = 5; i >= 0; --i) { FxTerm x = fx_term(t[i]); if (x.limb >= 0) b[x.limb] += x.val; }
          for (int l = 0; l < kFxLimbs; ++l) if (a[l] != b[l]) return 1;
          if (fx_value(a) != 7.25f) return 2;
          long long c[kFxLimbs] = {0};
          FxTerm x = fx_term(3e-30f); c[x.limb] += x.val; x = fx_term(-2e-30f); c[x.limb] += x.val;
          float want = (float)((double)3e-30f + (double)-2e-30f);
          if (fx_value(c) != want) return 3;
          x = fx_term(1e-45f); long long d[kFxLimbs] = {0}; d[x.limb] += x.val; if (fx_value(d) != 1e-45f) return 4;
          return 0;
        }''') % os.path.join(os.path.dirname(__file__), "..", "if-defense_b200", "csrc", "ifd_math.cuh")
    with tempfile.TemporaryDirectory() as td:
        f = os.path.join(td, "t.cpp")
        open(f, "w").write(src)
        subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-o", os.path.join(td, "t"), f])
        assert subprocess.call([os.path.join(td, "t")]) == 0


def test_loop_tracks_reference(mathcheck, conv):
    """20 Adam steps from the fixture's init points: the north_star tolerance (1e-4 max-abs) with margin."""
    pl, blob = planes_cl(conv), conv["dec_blob"]
    xyz = np.ascontiguousarray(conv["p0"]).copy()
    B, K, _ = xyz.shape
    steps = np.array([0, 1, 9, 19], np.int32)
    trace = np.zeros((4, B, K, 3), np.float32)
    mathcheck.mc_convonet_opt(P(blob), P(pl), P(xyz), B, K, 64, 5, 20, B, 5, D(1e-3), D(0.9), D(0.999), D(1e-8), D(0.2), D(500.),
                              D(0.07), D(0.03), D(1e-12), D(0.1), 1, P(steps), 4, P(trace))
    # rounding differences grow along the trajectory (DESIGN.md "Parity tiers"): tight early, 1e-4 at the end
    for n, (i, tol) in enumerate(zip(steps, (1e-6, 1e-6, 5e-6, 2e-5))):
        assert np.abs(trace[n] - conv["trace/xyz_%d" % i]).max() < tol, i
    assert np.abs(xyz - conv["final_20_normalized"]).max() < 1e-4


def test_late_state_single_step(mathcheck, conv):
    """Teacher-forced step 151 from the reference's own (xyz, m, v) at step 150: bias corrections at large t."""
    pl, blob = planes_cl(conv), conv["dec_blob"]
    xyz = np.ascontiguousarray(conv["trace/late_xyz"]).copy()
    m, v = conv["trace/late_m"].copy(), conv["trace/late_v"].copy()
    B, K, _ = xyz.shape
    g = np.zeros((B, K, 3), np.float32)
    mathcheck.mc_convonet_decode(P(blob), P(pl), P(xyz), B, K, 64, 5, D(0.1), 2, None, F32(0.2), F32(1.0 / B), None, P(g))
    gr = np.zeros((B, K, 3), np.float32)
    gl = np.full(B, 500.0 / B, np.float32)
    mathcheck.mc_knn_repulsion(P(xyz), B, K, 5, F32(0.07), F32(0.03), F32(1e-12), P(gl), None, None, P(gr), None)
    g = (g + gr).astype(np.float32)
    mathcheck.mc_adam(P(xyz), P(m), P(v), P(g), B * K * 3, D(1e-3), D(0.9), D(0.999), D(1e-8), 151)
    assert np.abs(xyz - conv["trace/late_xyz_next"]).max() < 1e-6
