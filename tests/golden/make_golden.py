#!/usr/bin/env python
"""Regenerate the committed golden fixtures from the reference tree.

Run here (the build container), never on the GPU box:

    python tests/golden/make_golden.py

It imports the reference's own ``src/utils/preprocessing.py`` (numpy / pandas /
sklearn only, so it runs unmodified) to reproduce the data split and the fitted
``StandardScaler``; reads the Keras HDF5 weights with the product's ``h5lite``
reader; parses the TF float32 predictions printed in
``tests/Xaccelerometer_quality_factor/surrogate_fft.ipynb`` cell 11 (pin K1);
and evaluates the problem set-up of the identification notebooks
(``identification.ipynb`` cells 6-14) with the numpy oracle.  Everything is
written as small ``.npz`` / ``.json`` files next to this script so that the GPU
tests, ``smoke()`` and ``bench.py`` never need ``/root/reference``.
"""
from __future__ import annotations

import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MEMS_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "mcmc-mems_b200"))
sys.path.insert(0, os.path.join(REF, "src", "utils"))

import h5lite  # noqa: E402
from oracle import make_forward_model  # noqa: E402
from preprocessing import preprocessing  # noqa: E402  (the reference's module)
from scipy.optimize import least_squares  # noqa: E402


def layers_of(path):
    return [(l.kernel, l.bias, l.activation) for l in h5lite.load_keras_dense_stack(path)]


def pack_layers(layers):
    out = {"n_layers": np.int64(len(layers)),
           "activations": np.array([a for _, _, a in layers])}
    for i, (k, b, _) in enumerate(layers):
        out[f"W{i}"] = k
        out[f"b{i}"] = b
    return out


def problem(test_dir, sigma2, low, high, starts, cov_factor, out_name):
    """identification.ipynb cells 6-14: data, model, x_true/y_true, y_obs, LSQ, covariance."""
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "tests", test_dir))
    try:
        dp = preprocessing("./config_I.json")
        layers = layers_of(dp.config["MODEL_PATH"])
    finally:
        os.chdir(cwd)
    sample = 10                                     # identification.ipynb cell 4
    x_true, y_true = dp.X_test[sample], dp.y_test[sample]
    T = len(dp.time)
    fwd = make_forward_model(dp.time, dp.scaler.mean_, dp.scaler.scale_, layers)
    # the notebooks draw y_observed unseeded (cell 10); SURVEY.md §8d fixes the seed
    y_obs = y_true + np.sqrt(sigma2) * np.random.default_rng(0).standard_normal(T)

    def resid(x):                                   # src/InverseProblems/utils.py:90-93
        return fwd(x).reshape(y_obs.shape) - y_obs

    x_opts, cov = [], None
    for sp in starts:                               # utils.py:83-88; cell 12 overwrites cov each pass
        res = least_squares(resid, sp, bounds=(low, high), jac="3-point")
        x_opts.append(res.x)
        cov = np.linalg.inv(res.jac.T @ res.jac)
    cov = cov_factor * cov
    np.savez(os.path.join(HERE, out_name),
             time=dp.time, scaler_mean=dp.scaler.mean_, scaler_scale=dp.scaler.scale_,
             x_true=x_true, y_true=y_true, y_obs=y_obs, sigma2=np.float64(sigma2),
             low=np.asarray(low, np.float64), high=np.asarray(high, np.float64),
             starts=np.asarray(starts, np.float64), x_opts=np.asarray(x_opts), cov=cov,
             X_test=dp.X_test[:32], y_test=dp.y_test[:32],
             f_true=fwd(x_true), **pack_layers(layers))
    return dp


def k1_fft():
    """Pin K1: TF float32 outputs printed in surrogate_fft.ipynb cell 11."""
    tdir = os.path.join(REF, "tests", "Xaccelerometer_quality_factor")
    cwd = os.getcwd()
    os.chdir(tdir)
    try:
        dp = preprocessing("./config_I_fft.json")
        layers = layers_of(dp.config["MODEL_PATH"])
    finally:
        os.chdir(cwd)
    nb = json.load(open(os.path.join(tdir, "surrogate_fft.ipynb")))
    cell = nb["cells"][11]
    text = "".join("".join(o["text"]) for o in cell["outputs"]
                   if o.get("output_type") == "stream" and o.get("name") == "stdout")
    arrays = [np.array(m.split(), dtype=np.float64) for m in re.findall(r"\[([^\]]*)\]", text)]
    arrays = [a for a in arrays if a.size == 15]
    assert len(arrays) == 20, len(arrays)
    tf_pred = np.stack(arrays[0::2])                # y_pred rows (TF float32)
    y_test = np.stack(arrays[1::2])                 # data rows printed next to them
    assert np.allclose(y_test, dp.y_test[:10], rtol=0, atol=1e-6)
    np.savez(os.path.join(HERE, "k1_fft.npz"), X_test_scaled=dp.X_test_scaled[:10],
             tf_pred=tf_pred, y_test=y_test, **pack_layers(layers))


def sensitivity():
    """Sensitivity surrogate of the geometric test (config_II.json: 3 -> 32x4 tanh -> 1) with its data split and
    scaler -- the inputs of plot_sensitivity_histogram (src/InverseProblems/plotsPaper.py:201-210; call site
    tests/Xaccelerometer_geometric/identification.ipynb:586-588)."""
    tdir = os.path.join(REF, "tests", "Xaccelerometer_geometric")
    cwd = os.getcwd()
    os.chdir(tdir)
    try:
        dp = preprocessing("./config_II.json")
        layers = layers_of(dp.config["MODEL_PATH"])
    finally:
        os.chdir(cwd)
    np.savez(os.path.join(HERE, "sensitivity.npz"), scaler_mean=dp.scaler.mean_, scaler_scale=dp.scaler.scale_,
             X_test=dp.X_test[:64], y_test=np.asarray(dp.y_test[:64], np.float64).reshape(-1),
             X_test_scaled=dp.X_test_scaled[:64], **pack_layers(layers))


def fft_coarse():
    """The FFT surrogate of the Q-factor test with the scaler of its own data processor (config_I_fft.json): the
    coarse first-stage model of delayed acceptance (oracle/coarse.py)."""
    tdir = os.path.join(REF, "tests", "Xaccelerometer_quality_factor")
    cwd = os.getcwd()
    os.chdir(tdir)
    try:
        dp = preprocessing("./config_I_fft.json")
        layers = layers_of(dp.config["MODEL_PATH"])
    finally:
        os.chdir(cwd)
    np.savez(os.path.join(HERE, "fft_coarse.npz"), scaler_mean=dp.scaler.mean_, scaler_scale=dp.scaler.scale_,
             X_test=dp.X_test[:16], y_test=dp.y_test[:16], y_series_test=dp.y_series_test[:16], **pack_layers(layers))


def main():
    if "--sensitivity-only" in sys.argv:
        sensitivity()
        return
    if "--fft-coarse-only" in sys.argv:
        fft_coarse()
        return
    sensitivity()
    fft_coarse()
    geo_bounds = ([0.1, -0.5, 29.0], [0.5, 0.5, 31.0])          # geometric cell 6
    geo_starts = [[0.3, 0.0, 30.0], [0.4, 0.25, 30.0], [0.2, 0.25, 30.0],
                  [0.4, -0.25, 30.0], [0.2, -0.25, 30.0]]
    problem("Xaccelerometer_geometric", (1e-6 * 1000 * np.sqrt(200) * 5) ** 2,
            *geo_bounds, geo_starts, 1.0, "geometric.npz")
    nbq = json.load(open(os.path.join(REF, "tests", "Xaccelerometer_quality_factor",
                                      "identification.ipynb")))
    src = "".join("".join(c["source"]) for c in nbq["cells"] if c["cell_type"] == "code")
    assert "10*cov_matrix" in src                               # Q-factor proposal is 10*cov
    qf_bounds = ([0.1, -0.5, 29.0, 0.4], [0.5, 0.5, 31.0, 0.6])
    qf_starts = [[0.3, 0.0, 30.0, 0.5], [0.4, 0.25, 30.0, 0.5], [0.2, 0.25, 30.0, 0.5],
                 [0.4, -0.25, 30.0, 0.5], [0.2, -0.25, 30.0, 0.5]]
    problem("Xaccelerometer_quality_factor", 0.5, *qf_bounds, qf_starts, 10.0, "qfactor.npz")
    k1_fft()
    pins = {}
    g = np.load(os.path.join(REF, "tests", "Xaccelerometer_geometric", "samples", "sample_10.npy"))
    q = np.load(os.path.join(REF, "tests", "Xaccelerometer_quality_factor", "samples", "sample_10.npy"))
    pins["geometric_sample_10"] = {"shape": list(g.shape), "mean": g.mean(1).tolist(), "std": g.std(1).tolist()}
    pins["qfactor_sample_10"] = {"shape": list(q.shape), "mean": q.mean(1).tolist(), "std": q.std(1).tolist()}
    # printed by the notebooks (identification.ipynb:254-282, Q-factor :478-506)
    pins["geometric_printed"] = {"acc": [0.276, 0.27466666666666667, 0.2795, 0.2693333333333333, 0.2728333333333333],
                                 "scale": [0.1314504627358983, 0.12692327229893097, 0.12892139843753464,
                                           0.12530442202049047, 0.1256356121192231],
                                 "rhat": [1.00452344, 1.00266526, 1.00371555]}
    pins["qfactor_printed"] = {"acc": [0.3486666666666667, 0.3541666666666667, 0.35283333333333333, 0.35733333333333334,
                                       0.3546666666666667],
                               "scale": [0.20871315242186197, 0.2091576605505331, 0.21424495688396644, 0.22466504907059892,
                                         0.21541593183414806],
                               "rhat": [1.00353547, 1.00247405, 1.00237344, 1.00205655]}
    json.dump(pins, open(os.path.join(HERE, "pins.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
