"""The C-ABI library loads on a CPU-only box and exports exactly what include/sinddm_b200.h declares.
No compute is launched here."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

REPO = Path(__file__).resolve().parents[1]
HEADER = REPO / "include" / "sinddm_b200.h"


def declared_functions():
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:int|size_t|void|const char\*|unsigned long long)\s+(sinddm_\w+)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    assert len(names) >= 24
    for must in ("sinddm_init", "sinddm_net_forward", "sinddm_net_backward", "sinddm_conv_forward",
                 "sinddm_conv_wgrad", "sinddm_dw5x5", "sinddm_ddpm_step", "sinddm_qsample_mix", "sinddm_l1_loss"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from sinddm_b200 import _capi
    lib = _capi.load()                      # raises if the .so is missing: there is no fallback
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _capi.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_capi.SIGNATURES) == declared_functions()
    assert lib.sinddm_abi_version() == 3


def test_workspace_queries_run_without_a_gpu():
    from sinddm_b200 import _capi
    lib = _capi.load()
    infer = lib.sinddm_plan_workspace_bytes(16, 186, 248, 160, 3, _capi.MATH_TF32, 0)
    train = lib.sinddm_plan_workspace_bytes(32, 186, 248, 160, 3, _capi.MATH_TF32, 1)
    assert 0 < infer < train < 40 * 2**30
    assert lib.sinddm_plan_workspace_bytes(0, 1, 1, 160, 3, 1, 0) == 0
    assert lib.sinddm_l1_loss_workspace_bytes() > 0
    assert lib.sinddm_conv_wgrad_workspace_bytes(2, 19, 23, 80, 80, 9, 1) > 0


def test_product_path_refuses_cpu_tensors():
    """No CPU fallback: CPU tensors (or a missing device) are an error, never a silent slow path."""
    from sinddm_b200 import SinDDMNet, ops, _capi
    net = SinDDMNet(dim=16, multiscale=True)
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 8, 8), torch.zeros(1, dtype=torch.long), 0)
    with pytest.raises(RuntimeError):
        ops.colsum(torch.zeros(4, 4))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            _capi.init(0)


def test_product_never_imports_the_oracle():
    for py in (REPO / "sinddm_b200").rglob("*.py"):
        src = py.read_text()
        assert "import oracle" not in src and "from oracle" not in src, py
    for py in list((REPO / "SinDDM").glob("*.py")) + [REPO / "main.py"]:
        if py.exists():
            src = py.read_text()
            assert "oracle" not in src, py


def test_math_modes_agree_between_header_binding_and_cli(monkeypatch):
    """SINDDM_MATH_* in the header == the ctypes constants == what `--math` / SINDDM_MATH select; a tf32x3 plan needs
    the split-operand buffers and the tripled weight operands on top of the TF32 plan's workspace."""
    import main as cli
    from sinddm_b200 import _capi, denoiser
    text = HEADER.read_text()
    m = re.search(r"enum \{ SINDDM_MATH_FP32 = (\d+), SINDDM_MATH_TF32 = (\d+), SINDDM_MATH_TF32X3 = (\d+) \};", text)
    assert m, "math-mode enum not found in the header"
    assert tuple(int(v) for v in m.groups()) == (_capi.MATH_FP32, _capi.MATH_TF32, _capi.MATH_TF32X3) == (0, 1, 2)
    for name, val in (("fp32", 0), ("tf32", 1), ("tf32x3", 2)):
        monkeypatch.setenv("SINDDM_MATH", name)
        assert denoiser.default_math() == val
        assert cli.build_parser().parse_args(["--math", name]).math == name
    monkeypatch.delenv("SINDDM_MATH")
    assert denoiser.default_math() == _capi.MATH_TF32          # the reference's own GPU default (cudnn.allow_tf32)
    lib = _capi.load()
    B, H, W, dim = 4, 94, 126, 160
    tf32 = lib.sinddm_plan_workspace_bytes(B, H, W, dim, 3, _capi.MATH_TF32, 1)
    x3 = lib.sinddm_plan_workspace_bytes(B, H, W, dim, 3, _capi.MATH_TF32X3, 1)
    split = 2 * 3 * B * H * W * dim * 4                          # split_a + split_b
    assert split <= x3 - tf32 <= split + 64 * 2**20              # + the [lo | hi | hi] weight operands (a few MB)


def test_conv_launch_picks_the_smallest_covering_epilogue_flavour():
    """tc_conv_kernel is instantiated per set of epilogue features; the launch picks the smallest instantiation that
    covers the descriptor (host-only query, no GPU).  The masks are the ones ncu shows in the kernel names."""
    from sinddm_b200 import _capi
    lib = _capi.load()
    P = 0x1000      # any non-NULL address: the query never dereferences

    def flavour(**kw):
        d = _capi.ConvDesc()
        d.B, d.H, d.W, d.Cin, d.N, d.ntaps = 1, 8, 8, 80, 80, 9
        d.inp, d.w, d.out = P, P, P
        for k, v in kw.items():
            setattr(d, k, v)
        return lib.sinddm_conv_epilogue_flavour(ctypes.byref(d))

    STREAM, PRE2, RESADD, DGELU, X3, GELU, PRE, FINAL, OUT3, COLSUM, ROUND = (1 << i for i in range(11))
    assert flavour() == 0                                                   # residual-slice layers, plain data gradients
    assert flavour(bias=P) == 0
    assert flavour(gelu=1) == flavour(gelu=1, out_pre=P, round_tf32=1) == GELU | PRE | ROUND          # net[0]
    assert flavour(x3=P, w_res3=P) == X3                                    # l1.net[2]
    assert flavour(res_add=P) == STREAM | RESADD                            # l3.net[2]
    assert flavour(w_final=P, b_final=P, out_final=P) == FINAL              # l4.net[2]
    assert flavour(dgelu_z=P, round_tf32=1) == STREAM | DGELU | COLSUM | ROUND      # net[2] data gradient
    assert flavour(round_tf32=1) == GELU | PRE | ROUND                      # smallest instantiation that can round
    assert flavour(res_add=P, dgelu_z=P) == 2047                            # two streamed operands: only the generic one
    assert flavour(gelu=1, res_add=P) == 2047
    assert lib.sinddm_conv_epilogue_flavour(None) < 0
