"""SURVEY.md §8 row a8 — the reference's own callers run unchanged on the drop-in layers.

* CPU, build container only (needs ``/root/reference``): the rebinding snippet of INTEGRATION.md §1 is extracted from
  the document and executed verbatim in a fresh interpreter; the reference's unmodified
  ``modules.my_models_graph.UNetSpherical`` must then build on the dsw layer classes with the very same state-dict keys
  and shapes as the pure reference model, and load its checkpoint with ``strict=True`` (``utils_config.py:409-413``).
* GPU: the reference's ResBlock tail exactly as written in ``my_models_graph.py:205-216`` — in-place ``*=`` / ``+=`` on
  the convolution's output, a ``torch.nn.Linear`` skip, a separate ``F.relu`` — on the CUDA ``ConvCheb`` (no fused
  activation, no fused ReZero tail, no NodeLinear), against the golden vectors of the unmodified reference.
"""
import os
import re
import subprocess
import sys
import textwrap
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from _util import REL_TOL, coo_from, golden, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("DSW_REFERENCE_ROOT", "/root/reference")


def _integration_snippet() -> str:
    with open(os.path.join(ROOT, "INTEGRATION.md")) as f:
        text = f.read()
    section = text.split("## 1.", 1)[1]
    return re.search(r"```python\n(.*?)```", section, flags=re.S).group(1)


@pytest.mark.skipif(not os.path.isfile(os.path.join(REFERENCE, "modules", "layers.py")),
                    reason="the unmodified reference is only present in the build container")
@pytest.mark.parametrize("pool_method", ["interp", "max", "maxval"])
def test_reference_unet_builds_on_rebound_layers(pool_method, lib):
    script = textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r})
        import torch
        from oracle.ref_import import load_reference
        from deepsphere_weather_b200 import graphs as G, models as M
        ref_layers, ref_models = load_reference()
        ref_layers.build_pooling_matrices = lambda src, dst: G.nested_pool_matrices(src.n_vertices, 4)  # CDO is absent
        ti = M.default_tensor_info(768)
        kw = dict(kernel_size_conv=4, pool_method={pool_method!r})
        torch.manual_seed(0)
        pure = ref_models.UNetSpherical(ti, "healpix", {{"subdivisions": 8, "nest": True}}, **kw)
        assert type(pure.conv1.convblock1.conv).__module__ == "modules.layers"
        # ---- INTEGRATION.md section 1, verbatim ----
        SNIPPET
        # --------------------------------------------
        rebound = ref_models.UNetSpherical(ti, "healpix", {{"subdivisions": 8, "nest": True}}, **kw)
        conv = rebound.conv1.convblock1.conv
        assert type(conv).__module__.startswith("deepsphere_weather_b200"), type(conv)
        assert type(rebound.pool1).__module__.startswith("deepsphere_weather_b200"), type(rebound.pool1)
        sd_ref, sd_new = pure.state_dict(), rebound.state_dict()
        assert list(sd_ref) == list(sd_new), set(sd_ref) ^ set(sd_new)
        for k in sd_ref:
            assert tuple(sd_ref[k].shape) == tuple(sd_new[k].shape), k
            assert sd_ref[k].is_sparse == sd_new[k].is_sparse, k
        rebound.load_state_dict(sd_ref, strict=True)     # a reference checkpoint loads into the rebound model
        pure.load_state_dict(rebound.state_dict(), strict=True)   # ... and the other way round
        for k, v in rebound.state_dict().items():
            if not v.is_sparse:
                assert torch.equal(v, sd_ref[k]), k
        print("OK", len(sd_ref))
    """).replace("SNIPPET", _integration_snippet())
    env = dict(os.environ, PYTHONPATH=REFERENCE + os.pathsep + ROOT)
    out = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().splitlines()[-1].startswith("OK")


@pytest.mark.gpu
@pytest.mark.parametrize("mix_mode", [0, 1], ids=["mix-fp32", "mix-tcgen05"])
def test_reference_inplace_resblock_tail_on_cuda_convcheb(mix_mode, lib):
    """``x_out *= rezero_weight; x_out += res_connection(x)`` with ``torch.nn.Linear`` and a separate ``F.relu``
    (``my_models_graph.py:104-118, 205-216``) acting in place on the CUDA convolution's output."""
    from deepsphere_weather_b200 import layers as L
    from deepsphere_weather_b200 import models as M
    from oracle.unet_oracle import fill_parameters

    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests (there is no CPU fallback)")
    dev = torch.device("cuda:0")

    class PlainConvCheb(L.ConvCheb):
        """The reference's call signature only: ``forward(inputs)``, nothing fused."""
        fused_activations = ()

        def forward(self, inputs):
            return super().forward(inputs)

    backend = SimpleNamespace(
        ConvCheb=PlainConvCheb,
        healpix_pools={"max": (L.HealpixMaxPool, L.HealpixMaxUnpool), "avg": (L.HealpixAvgPool, L.HealpixAvgUnpool)},
        general_pools=L.PoolUnpoolBlock.getGeneralPoolUnpoolLayer,
    )  # no Linear, no rezero_residual: ResBlock falls back to torch.nn.Linear and the in-place tail
    g = golden("unet_interp_k4")
    laps = [coo_from(g, f"lap{i}") for i in range(3)]
    prev = lib.dsw_get_mix_mode()
    lib.dsw_set_mix_mode(mix_mode)
    try:
        model = M.UNetSpherical(M.default_tensor_info(768), "healpix", {"subdivisions": 8, "nest": True},
                                kernel_size_conv=4, pool_method="interp", laplacians=laps, backend=backend)
        assert isinstance(model.conv1.res_connection, torch.nn.Linear) and not isinstance(model.conv1.res_connection, L.NodeLinear)
        assert model.conv1._fused_tail is None and model.conv1.convblock1._fused_act is None
        fill_parameters(model, 11)
        model = model.to(dev)
        launches0 = lib.dsw_launch_count()
        y = model(torch.from_numpy(g["x"]).to(dev))
        assert lib.dsw_launch_count() > launches0
        assert rel_err(y, g["y"]) < REL_TOL
        loss = (y**2).mean()
        loss.backward()
        assert abs(loss.item() - float(g["loss"])) < REL_TOL * abs(float(g["loss"]))
        grads = dict(model.named_parameters())
        for n, ref_norm in zip([str(n) for n in g["grad_names"]], g["grad_norms"]):
            got = grads[n].grad.norm().item()
            assert abs(got - ref_norm) <= 1e-3 * max(ref_norm, 1e-6) + 1e-9, n
        for key in g.files:
            if key.startswith("grad__"):
                assert rel_err(grads[key[6:]].grad, g[key]) < 2 * REL_TOL, key
    finally:
        lib.dsw_set_mix_mode(prev)
