"""CPU: the oracle restatement (oracle/) against the golden vectors frozen from the
UNMODIFIED reference code (tests/golden/make_golden.py).  fp32 CPU on both sides, so
tolerances are tight (summation-order noise only)."""
import numpy as np
import torch

from helpers import load_golden, oracle_model_from_golden, t, WAVELENGTHS
from oracle import nerf3p, umhs_oracle as O

RTOL, ATOL = 1e-5, 1e-6


def close(a, b, rtol=RTOL, atol=ATOL):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(a.reshape(np.asarray(b).shape), b, rtol=rtol, atol=atol)


def test_colour_cluster_dense_weights():
    d, _ = load_golden("colour.npz")
    cs = O.ColourSystem(d["wavelengths"].tolist())
    close(cs.transform_matrix, d["transform_matrix"], 1e-6, 1e-8)
    close(cs(t(d["spec"])), d["rgb"])
    inner, probs = O.cluster_lookup(t(d["spec"]), 0.2, t(d["endmembers"]))
    close(inner, d["inner"])
    close(probs, d["probs"])
    w = O.get_weights_spectral(t(d["deltas"]), t(d["densities"]))
    close(w, d["weights"])
    close(O.spectral_render(t(d["vals"]), w), d["dense_acc"])


def _field_case(name):
    d, meta = load_golden(name)
    m = oracle_model_from_golden(d, meta)
    f = m.field
    rs = nerf3p.RaySamples(nerf3p.Frustums(t(d["origins"]), t(d["directions"]), t(d["starts"]), t(d["ends"])))
    density, emb = f.get_density(rs)
    out = f.get_outputs(rs, emb)
    close(density, d["density"])
    close(emb, d["embedding"])
    close(out["spectral"], d["spectral"])
    close(out["abundances"], d["abundances"])
    N = meta["N"]
    loss = (out["spectral"].reshape(N, 128) * t(d["gw"])).sum() + (density * t(d["gd"])).sum()
    if meta["specular"]:
        close(out["spectral2"], d["spectral2"])
        close(out["specular"], d["specular"])
        loss = loss + 0.5 * (out["spectral2"] * t(d["gw"]).flip(0)).sum()
    close(loss, d["loss"], 1e-5, 1e-4)
    loss.backward()
    n = 0
    for k, p in f.named_parameters():
        if "grad." + k in d:
            close(p.grad, d["grad." + k], 1e-4, 1e-5)
            n += 1
    assert n >= 10


def test_field_k4():
    _field_case("field_k4.npz")


def test_field_k6_specular():
    _field_case("field_k6_spec.npz")


def _model_case(name):
    d, meta = load_golden(name)
    training = meta["training"]
    m = oracle_model_from_golden(d, meta, training)
    assert abs(m.config.render_step_size - meta["render_step_size"]) < 1e-12
    rb = nerf3p.RayBundle(t(d["origins"]), t(d["directions"]))
    out = m.get_outputs(rb, jitter=t(d["jitter"]))
    # packing: bit exact
    assert np.array_equal(out["ray_indices"].numpy(), d["ray_indices"])
    assert np.array_equal(out["t_starts"].numpy(), d["t_starts"])
    assert np.array_equal(out["t_ends"].numpy(), d["t_ends"])
    for k in ("accumulation", "depth", "spectral", "rgb", "abundances", "seg_probs", "spectral2", "specular"):
        if "out." + k in d:
            close(out[k], d["out." + k], 1e-4, 1e-6)
    assert np.array_equal(out["num_samples_per_ray"].numpy(), d["out.num_samples_per_ray"])
    assert np.array_equal(out["seg_raw"].numpy(), d["out.seg_raw"])
    loss = m.get_loss_dict(out, {"image": t(d["image"]), "hs_image": t(d["hs_image"])}, bg=t(d["bg"]))
    for k, v in loss.items():
        close(v, d["loss." + k], 1e-5, 1e-7)
    if training:
        sum(loss.values()).backward()
        for k, p in m.field.named_parameters():
            if "grad." + k in d:
                close(p.grad, d["grad." + k], 2e-4, 1e-7)


def test_model_train_k4():
    _model_case("model_train_k4.npz")


def test_model_train_k4_specular():
    _model_case("model_train_k4_spec.npz")


def test_model_eval_k9_specular():
    _model_case("model_eval_k9_spec.npz")


def test_pack_oracle_properties():
    """Size-independent properties of the march: sorted, non-overlapping, inside [near, far]."""
    d, meta = load_golden("model_train_k4.npz")
    ri, t0, t1 = d["ray_indices"], d["t_starts"], d["t_ends"]
    assert (np.diff(ri) >= 0).all()
    same = np.diff(ri) == 0
    assert (t1 > t0).all()
    assert (t0[1:][same] >= t1[:-1][same] - 1e-6).all()
    assert t0.min() >= 0.05
