"""Pins the oracle against the LIVE reference at its real channel widths (BASELINE config 1: resolution-128
config, depth 1 and 2, batch 4, logistic+R1, EMA).  Only runs where /root/reference exists (the build container);
the reference is executed in a subprocess because its top-level package is also called `models`."""
import os
import subprocess
import sys

import pytest
import torch

import stylegan_oracle as O
from test_oracle_golden import arch_from, rel

REF = os.environ.get("SG_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference not mounted")
def test_oracle_matches_live_reference_full_width(tmp_path):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "make_golden.py"), "--live", str(tmp_path)],
                   check=True, env=env, timeout=900)
    fx = torch.load(tmp_path / "live_full128.pt", weights_only=False)
    arch = arch_from(fx["arch"])
    st = O.TrainState(arch, fx["g_init"], fx["d_init"], use_ema=True, ema_decay=0.999)
    for (depth, alpha), s in zip(fx["schedule"], fx["steps"]):
        nz = fx["noises"][:2 * (depth + 1)]
        dl, gl = O.train_step(st, s["images"], s["z"], depth, alpha, s["mix_d"], s["mix_g"], nz, nz)
        assert abs(dl - s["d_loss"]) <= 1e-3 * abs(s["d_loss"]), (depth, alpha, dl, s["d_loss"])
        assert abs(gl - s["g_loss"]) <= 1e-3 * abs(s["g_loss"]), (depth, alpha, gl, s["g_loss"])
    for k, v in fx["d_final"].items():
        assert rel(st.d[k], v) < 2e-2, k
    for k, v in fx["g_final"].items():
        assert rel(st.g[k], v) < 2e-2, k
