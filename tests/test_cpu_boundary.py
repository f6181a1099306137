"""CPU (no GPU needed): the C-ABI library builds for sm_100a, loads, exports every symbol that
include/pcm_b200.h declares, and the product path refuses to run without the GPU / the library."""
import ctypes
import os
import re

import pytest
import torch

import pcm_b200
from pcm_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    return pcm_b200.build()


def test_header_symbols_exported(built):
    dll = ctypes.CDLL(built)
    with open(os.path.join(ROOT, "include", "pcm_b200.h")) as f:
        declared = set(re.findall(r"\b(pcm_\w+)\s*\(", f.read()))
    assert len(declared) >= 30
    assert declared == set(_lib.PROTOS), declared ^ set(_lib.PROTOS)
    for name in declared:
        assert hasattr(dll, name), f"{name} declared in the header but not exported by libpcm_b200.so"


def test_version_and_error_string_without_gpu(built):
    L = _lib.lib()
    assert L.version() >= 100
    assert isinstance(L.last_error(), str)


def test_library_is_sm100a_only(built):
    out = os.popen(f"cuobjdump -lelf {built} 2>/dev/null").read()
    if not out.strip():
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out and not re.search(r"sm_(?!100a)\d+", out), out


def test_cpu_tensors_are_refused():
    from pcm_b200 import ops
    from pcm_b200.src.unet import ConvBlock
    from pcm_b200.src.unet_convlstm_attention import AttUNetConvLSTM
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ConvBlock(8, 16)(torch.zeros(1, 8, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        AttUNetConvLSTM(7, 2, 8)(torch.zeros(1, 2, 7, 16, 24))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.mse_loss(torch.zeros(4), torch.zeros(4))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_build, "LIB_PATH", str(tmp_path / "libpcm_b200.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib._Lib()


def test_module_surface_matches_reference_appendix_c():
    """state_dict keys/shapes of the drop-in modules == the oracle's spec (which make_goldens.py loaded into
    the REAL reference modules with strict=True)."""
    from oracle import model_oracle as O
    from pcm_b200.src.models import get_model
    from pcm_b200.src.unet import UNet
    from pcm_b200.src.unet_convlstm_attention import AttUNetConvLSTM
    for mod, spec in [(AttUNetConvLSTM(7, 2, 16), O.attunet_spec(7, 2, 16)), (UNet(5, 2, 16), O.unet_spec(5, 2, 16))]:
        sd = mod.state_dict()
        assert list(sd.keys()) == [k for k, _ in spec]
        for k, shape in spec:
            assert tuple(sd[k].shape) == tuple(shape), k
    cfg = {"model": {"type": "unet_convlstm_attention", "base_channels": 16},
           "data": {"input_vars": ["CO2", "SO2", "CH4", "BC", "rsdt"], "output_vars": ["tas", "pr"]}}
    m = get_model(cfg)
    assert isinstance(m, AttUNetConvLSTM) and m.enc1.body[0].weight.shape[1] == 7      # in_ch=7 hard-coded
    assert sum(p.numel() for p in m.parameters()) == 953968
    with pytest.raises(ValueError, match="Unknown model type"):
        get_model({"model": {"type": "nope"}, "data": cfg["data"]})


def test_default_init_matches_reference_under_seed():
    """Same registration order as the reference ctor => same default init under torch.manual_seed(42)
    (checksums recorded from the real reference by oracle/make_goldens.py)."""
    import json
    from pcm_b200.src.unet_convlstm_attention import AttUNetConvLSTM
    with open(os.path.join(ROOT, "tests", "golden", "metric_appendix_g.json")) as f:
        want = json.load(f)["attunet_default_init_seed42"]
    torch.manual_seed(42)
    m = AttUNetConvLSTM(7, 2, 16)
    for k, v in m.state_dict().items():
        s, n = float(v.double().sum()), float(v.double().norm())
        assert abs(s - want[k][0]) <= 1e-9 * max(1.0, abs(want[k][0])) and abs(n - want[k][1]) <= 1e-9 * max(1.0, want[k][1]), k


def test_parameter_surface_and_default_init_match_reference():
    """Same seed -> bit-identical state_dict (keys, order, shapes, values incl. BatchNorm buffers) as the reference's own
    constructors, for every model of the family (only where /root/reference is mounted, i.e. in the build container)."""
    import pytest
    import torch
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference sources not mounted")
    import pcm_b200  # noqa: F401
    R = ref_loader.load()
    from pcm_b200.src import models as M, unet as U
    from pcm_b200.src.cnn_transformer import CNNTransformer
    from pcm_b200.src.convlstm import ConvLSTM
    from pcm_b200.src.unet_convlstm_attention import AttUNetConvLSTM
    pairs = [
        (lambda: M.SimpleCNN(5, 2), lambda: R.models.SimpleCNN(5, 2)),
        (lambda: M.SimpleCNN(5, 2, init_dim=8, depth=2), lambda: R.models.SimpleCNN(5, 2, init_dim=8, depth=2)),
        (lambda: M.ResidualBlock(8, 8), lambda: R.models.ResidualBlock(8, 8)),
        (lambda: U.UNet(5, 2, 16), lambda: R.unet.UNet(5, 2, 16)),
        (lambda: AttUNetConvLSTM(7, 2, 16), lambda: R.unet_convlstm_attention.AttUNetConvLSTM(7, 2, 16)),
        (lambda: CNNTransformer(), lambda: R.cnn_transformer.CNNTransformer()),
        (lambda: ConvLSTM(16, 8), lambda: R.convlstm.ConvLSTM(16, 8)),
    ]
    for ours, ref in pairs:
        torch.manual_seed(42)
        a = ours().state_dict()
        torch.manual_seed(42)
        b = ref().state_dict()
        assert list(a.keys()) == list(b.keys())
        for k in a:
            assert torch.equal(a[k], b[k]), k


def test_every_kernel_follows_the_dependent_launch_protocol():
    """Kernels are launched with programmatic stream serialization (csrc/common.cuh): each __global__ function must
    execute griddepcontrol.wait (PCM_PDL_ENTRY or pdl_wait) and no launch may bypass pcm::launch."""
    import glob
    import os
    import re
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "physics-based-climate-model_b200", "csrc")
    n_kernels = 0
    for path in sorted(glob.glob(os.path.join(csrc, "*.cu"))):
        src = open(path).read()
        assert "<<<" not in src, f"{path}: raw <<< >>> launch (use pcm::launch)"
        for m in re.finditer(r"__global__", src):
            start = src.index("{", _skip_signature(src, m.end()))
            depth, i = 0, start
            while True:
                depth += src[i] == "{"
                depth -= src[i] == "}"
                if depth == 0:
                    break
                i += 1
            body = src[start:i]
            n_kernels += 1
            assert "PCM_PDL_ENTRY()" in body or ("pdl_wait()" in body and "pdl_launch_dependents()" in body), \
                f"{path}: kernel at offset {m.start()} has no griddepcontrol.wait"
            first_ret = body.find("return")
            first_wait = min(x for x in (body.find("PCM_PDL_ENTRY()"), body.find("pdl_wait()")) if x >= 0)
            assert first_ret < 0 or first_wait < first_ret, f"{path}: kernel at offset {m.start()} can return before the wait"
    assert n_kernels >= 50


def _skip_signature(src, pos):
    """index just past the parameter list of the kernel whose __global__ keyword ends at pos"""
    i = src.index("(", pos)
    while src[pos:i].rstrip().endswith("__launch_bounds__"):
        i = src.index("(", _close(src, i))
    return _close(src, i)


def _close(src, i):
    depth = 0
    while True:
        depth += src[i] == "("
        depth -= src[i] == ")"
        i += 1
        if depth == 0:
            return i


def test_product_synth_recipe_equals_the_oracle_recipe():
    """bench.py's product arm synthesises its inputs with pcm_b200.synth (no oracle import on that path); the oracle keeps
    its own statement of the same recipe for the checker side — both must produce identical bits."""
    import torch
    from oracle import model_oracle as O
    from pcm_b200 import synth
    a = synth.synth_attunet_batch(8, 3, 8, 12, seed=9)
    b = O.synth_attunet_batch(8, 3, 8, 12, seed=9)
    assert all(torch.equal(u, v) for u, v in zip(a, b))
    a = synth.synth_frame_batch(3, 5, 8, 12, seed=4)
    b = O.synth_frame_batch(3, 5, 8, 12, seed=4)
    assert all(torch.equal(u, v) for u, v in zip(a, b))
    import ast
    import os
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py")).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "run_ours")
    for node in ast.walk(fn):
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            mod = getattr(node, "module", None) or ""
            names = [a.name for a in node.names]
            assert not mod.startswith("oracle") and not any(n.startswith("oracle") for n in names), \
                "bench.py run_ours must not import oracle (only its baseline legs may)"


def test_gate_sum_handover_slot(built, monkeypatch):
    """ops._offer_sdot / ops._take_sdot (host logic of the pooling-backward -> block-backward hand-over): the sum is
    delivered only to the backward that receives exactly the gradient tensor it was formed for, the slot never outlives
    the next ConvBlock backward, and PCM_POOL_SDOT=0 / unsupported channel counts disable it."""
    from pcm_b200 import ops
    monkeypatch.delenv("PCM_POOL_SDOT", raising=False)
    ds = torch.zeros(2, 4, 6, 16)
    s = ops._offer_sdot(ds)
    assert s is not None and s.shape == (2 * 4 * 6,) and s.dtype == torch.float32
    assert ops._take_sdot(ds) is s                      # same tensor: delivered
    assert ops._take_sdot(ds) is None                   # ... once
    s = ops._offer_sdot(ds)
    other = torch.zeros(2, 4, 6, 16)
    assert ops._take_sdot(other) is None                # another gradient: not delivered, and the slot is emptied
    assert ops._take_sdot(ds) is None
    s = ops._offer_sdot(ds)
    assert ops._take_sdot(ds.view(2, 4, 6, 16)) is s    # a view with the same address / shape is the same data
    assert ops._offer_sdot(torch.zeros(1, 2, 2, 24)) is None          # C/8 = 3 is not a power of two
    assert ops._offer_sdot(torch.zeros(1, 2, 2, 12)) is None          # C % 8 != 0
    assert ops._take_sdot(ds) is None                   # a refused offer leaves the slot empty
    monkeypatch.setenv("PCM_POOL_SDOT", "0")
    assert ops._offer_sdot(ds) is None
    # the support query the offer consults on the device is host-only arithmetic (no GPU needed): the level-1 image of
    # config 3 fits one SM in bf16, the config-5 one does not
    L = _lib.lib()
    assert L._fn["pcm_convblock_fused_supported"](48, 72, 16, 2, 1) == 1
    assert L._fn["pcm_convblock_fused_supported"](184, 360, 32, 4, 1) == 0


def test_cost_model_of_the_gate_sum_entry_points(built):
    """costmodel.algo_cost (bench.py's roofline numerators) for the session-4 entry points: with the per-pixel gate sum
    supplied the backward tail's algorithm is dout + x in, dx out (+ 13 bytes of maps / ties and 4 bytes of sum per pixel);
    without it, it is pcm_convblock_tail_bwd's four tensors.  Argument order comes from the parsed header."""
    from pcm_b200 import costmodel
    names = [an for _, an in _lib.PROTOS["pcm_convblock_tail_bwd_sdot"][1]]
    vals = {"N": 384, "H": 48, "W": 72, "C": 16, "Cr": 2, "dtype": 1}
    px = 384 * 48 * 72

    def args(**over):
        return [over.get(n, vals.get(n, 1)) for n in names]                # non-zero stand-ins for pointers

    fl, by, bound = costmodel.algo_cost("pcm_convblock_tail_bwd_sdot", args(dq_out=0))
    assert (fl, bound) == (0.0, "hbm") and by == 3 * px * 16 * 2 + 17 * px
    _, by0, _ = costmodel.algo_cost("pcm_convblock_tail_bwd_sdot", args(dq_out=0, sdot=0))
    names_old = [an for _, an in _lib.PROTOS["pcm_convblock_tail_bwd"][1]]
    _, by_old, _ = costmodel.algo_cost("pcm_convblock_tail_bwd", [vals.get(n, 1) for n in names_old])
    assert by0 == by_old == 4 * px * 16 * 2 + 13 * px
    assert costmodel.shape_key("pcm_convblock_tail_bwd_sdot", args()) == "convblock_tail_bwd_sdot[384,48,72,16,2]"
    names_p = [an for _, an in _lib.PROTOS["pcm_maxpool2_bwd_skip_dot"][1]]
    vp = {"N": 384, "H": 48, "W": 72, "C": 16, "T": 6, "t_major": 1, "dtype": 1}
    _, byp, _ = costmodel.algo_cost("pcm_maxpool2_bwd_skip_dot", [vp.get(n, 1) for n in names_p])
    _, byq, _ = costmodel.algo_cost("pcm_maxpool2_bwd_skip_dot", [0 if n == "sdot" else vp.get(n, 1) for n in names_p])
    assert byp - byq == 4 * px
