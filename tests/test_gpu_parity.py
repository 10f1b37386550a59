"""Parity tests proper: the CUDA path (through cc_b200.* and the C ABI of libccb200.so) against the
oracle and the golden fixtures, on the B200."""
import pytest
import torch
from tests import kernel_cases as KC

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def cuda_lib():
    from cc_b200 import _lib, pyramid
    _lib._lib = None                      # make sure the real library (not a simulator) is bound
    assert not _lib.is_simulator(), 'GPU tests must run on the sm_100a library'
    pyramid.clear()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


@pytest.mark.parametrize('case', KC.ALL_CASES, ids=lambda f: f.__name__)
def test_case(case):
    case(torch.device('cuda:0'))
    torch.cuda.synchronize()


def test_full_size_rigid_loss_vs_cpu_oracle():
    """BASELINE.json size (b4, 256x832, 6 levels): product vs the CPU oracle; masks bit-exact."""
    KC.case_rigid_loss_oracle(torch.device('cuda:0'), B=4, H=256, W=832, NL=6, seed=5, robust=True,
                              oracle_device=torch.device('cpu'))
    KC.case_occlusion_and_valid_masks(torch.device('cuda:0'), B=4, H=256, W=832, NL=6, seed=5)


def test_full_size_properties():
    """Size-independent properties at full size: identity pose + huge depth => warp is the identity on
    interior pixels; zero flow samples at pixel centres - 0.5 (align_corners=False quirk)."""
    from cc_b200 import inverse_warp as CW, synth
    dev = torch.device('cuda:0')
    B, H, W = 4, 256, 832
    tgt, refs = synth.frames(B, H, W, seed=9)
    K, Kinv = synth.intrinsics(B, H, W)
    img, K, Kinv = refs[0].to(dev), K.to(dev), Kinv.to(dev)
    flow = torch.zeros(B, 2, H, W, device=dev)
    out = CW.flow_warp(img, flow)
    exp = torch.nn.functional.grid_sample(
        img, torch.stack(torch.meshgrid(torch.linspace(-1, 1, H, device=dev), torch.linspace(-1, 1, W, device=dev),
                                        indexing='ij')[::-1], -1)[None].expand(B, H, W, 2),
        align_corners=False)
    assert (out - exp).abs().max().item() < 1e-5
    # linearity of the backward in grad_out
    depth = synth.depths(B, H, W, 1, seed=2)[0][:, 0].to(dev).requires_grad_(True)
    pose = synth.poses(B, 4, seed=3)[:, 0].to(dev).requires_grad_(True)
    o = CW.inverse_warp(img, depth, pose, K, Kinv)
    g1 = torch.autograd.grad(o.sum(), [depth, pose], retain_graph=True)
    g2 = torch.autograd.grad((2.5 * o).sum(), [depth, pose])
    for a, b in zip(g1, g2):
        assert (2.5 * a - b).abs().max().item() <= 1e-5 * b.abs().max().item() + 1e-12


from tests import net_cases as NC   # noqa: E402


@pytest.mark.parametrize('case', NC.NET_CASES, ids=lambda f: f.__name__)
def test_net_case(case):
    case(torch.device('cuda:0'))
    torch.cuda.synchronize()


def test_conv_big_shapes():
    NC.case_conv_shapes(torch.device('cuda:0'), big=True)


from tests import step_cases as SC   # noqa: E402


def test_flat_adam():
    SC.case_flat_adam(torch.device('cuda:0'))


def test_train_step_cfg1_vs_oracle():
    from cc_b200 import nn as cnn, _lib
    saved = cnn.CONV_IMPL
    try:
        cnn.CONV_IMPL = _lib.IMPL_FFMA
        SC.case_step_cfg1(torch.device('cuda:0'), gtol=4e-3)
        cnn.CONV_IMPL = _lib.IMPL_AUTO
        SC.case_step_cfg1(torch.device('cuda:0'), gtol=5e-2)
    finally:
        cnn.CONV_IMPL = saved


def test_conv_tensor_core_path():
    NC.case_conv_tc(torch.device('cuda:0'))


def test_conv_tma_family():
    NC.case_conv_tma_family(torch.device('cuda:0'))


def test_alternate_nets():
    """All six alternate architectures against the fixtures frozen from the reference's modules."""
    NC.case_alt_nets(torch.device('cuda:0'))


def test_conv_nhwc_slab():
    NC.case_conv_nhwc(torch.device('cuda:0'))


def test_joint_step_cfg3_vs_oracle():
    SC.case_step_cfg3(torch.device('cuda:0'))


def test_full_size_loss_layer_vs_cpu_oracle():
    """The flow-photometric, smoothness, BCE and consensus kernels at the BASELINE size (b4 256x832, 6 levels: 960+
    CTAs through the tile/prefix tables) against the CPU oracle."""
    KC.case_loss_layer_fullsize(torch.device('cuda:0'), B=4, H=256, W=832, NL=6, oracle_device=torch.device('cpu'))
    KC.case_consensus_fullsize(torch.device('cuda:0'), B=4, H=256, W=832, NL=6)


from tests import io_cases as IC   # noqa: E402


@pytest.mark.parametrize('case', IC.IO_CASES, ids=lambda f: f.__name__)
def test_io_case(case):
    case(torch.device('cuda:0'))


def test_io_full_size():
    IC.case_metrics_oracle_sizes(torch.device('cuda:0'))
    IC.case_input_pipeline_fullsize(torch.device('cuda:0'))


from tests import helper_cases as HC   # noqa: E402


@pytest.mark.parametrize('case', HC.HELPER_CASES, ids=lambda f: f.__name__)
def test_helper_case(case):
    case(torch.device('cuda:0'))


from tests import eval_cases as EC   # noqa: E402


@pytest.mark.parametrize('case', EC.EVAL_CASES, ids=lambda f: f.__name__)
def test_eval_case(case):
    """Evaluation cores (test_disp / test_pose / test_flow sample loops) against the oracle restatement."""
    case(torch.device('cuda:0'))
