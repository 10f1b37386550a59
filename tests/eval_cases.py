"""Parity cases of the evaluation cores (cc_b200/evaluate.py) against the oracle restatement of the
reference's test_disp.py / test_pose.py / test_flow.py sample loops (oracle/evaluate.py) on synthetic samples:
the same uint8 frames and ground truth through both, nets loaded with the same state_dict."""
import numpy as np
import torch
from cc_b200 import synth, models as CM, evaluate as CE
from oracle import nets as ON, evaluate as OE
from tests.net_cases import _load


def _frames_u8(n, H, W, seed):
    tgt, refs = synth.frames(1, H, W, seed=seed)
    fr = [refs[0], refs[1], tgt, refs[2], refs[3]][:n] if n == 5 else [refs[1], tgt, refs[2]]
    return [np.clip((f[0].permute(1, 2, 0).numpy() * 0.5 + 0.5) * 255, 0, 255).astype(np.uint8) for f in fr]


def case_eval_depth(device):
    H, W = 64, 128
    imgs = _frames_u8(5, H, W, seed=71)
    tgt, refs = imgs[2], imgs[:2] + imgs[3:]
    rs = np.random.RandomState(5)
    gt = (2.0 + 30.0 * rs.rand(96, 200)).astype(np.float32)
    mask = rs.rand(96, 200) > 0.3
    disp_w, pose_w = ON.disp_params(), ON.pose_params()
    dnet = _load(CM.DispResNet6(), disp_w, device)
    pnet = _load(CM.PoseNetB6(nb_ref_imgs=4), pose_w, device)
    displacements = [0.8, 0.4, 0.0, 0.9]
    got = CE.depth_sample_errors(dnet, tgt, gt, mask, 1e-3, 80.0, pnet, refs, displacements, device=device)
    want = OE.depth_sample_errors(disp_w, tgt, gt, mask, 1e-3, 80.0, pose_w, refs, displacements)
    assert np.allclose(got, want, rtol=2e-3, atol=2e-4), (got, want)
    got2 = CE.depth_sample_errors(dnet, tgt, gt, None, 1e-3, 80.0, device=device)
    want2 = OE.depth_sample_errors(disp_w, tgt, gt, None, 1e-3, 80.0)
    assert np.allclose(got2, want2, rtol=2e-3, atol=2e-4) and (got2[0] == 0).all()


def case_eval_pose(device):
    H, W = 64, 128
    imgs = _frames_u8(5, H, W, seed=72)
    rs = np.random.RandomState(6)
    gt = np.zeros((5, 3, 4))
    for i in range(5):
        a = 0.02 * rs.randn(3)
        Rx = np.array([[1, 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
        Rz = np.array([[np.cos(a[2]), -np.sin(a[2]), 0], [np.sin(a[2]), np.cos(a[2]), 0], [0, 0, 1]])
        gt[i, :, :3] = Rx @ Rz
        gt[i, :, 3] = 0.3 * i + 0.05 * rs.randn(3)
    pose_w = ON.pose_params()
    pnet = _load(CM.PoseNetB6(nb_ref_imgs=4), pose_w, device)
    ate, re, final = CE.pose_snippet_errors(pnet, imgs, gt, device=device)
    ate_o, re_o, final_o = OE.pose_snippet_errors(pose_w, imgs, gt)
    assert np.allclose(final, final_o, rtol=1e-4, atol=1e-6)
    assert abs(ate - ate_o) <= 1e-4 * max(1.0, abs(ate_o)) and abs(re - re_o) <= 1e-4 * max(1.0, abs(re_o))


def case_eval_flow(device):
    H, W = 64, 128
    tgt, refs = synth.frames(1, H, W, seed=73)
    K, Kinv = synth.intrinsics(1, H, W)
    rs = np.random.RandomState(7)
    Hg, Wg = 96, 200
    flow_gt = torch.from_numpy(np.concatenate([3 * rs.randn(1, 2, Hg, Wg), (rs.rand(1, 1, Hg, Wg) > 0.2)], 1).astype(np.float32))
    obj = torch.from_numpy((rs.rand(1, Hg, Wg) > 0.7).astype(np.float32))
    P = dict(disp=ON.disp_params(), pose=ON.pose_params(), mask=ON.mask_params(), flow=ON.flow_params())
    nets = dict(disp=_load(CM.DispResNet6(), P['disp'], device), pose=_load(CM.PoseNetB6(nb_ref_imgs=4), P['pose'], device),
                mask=_load(CM.MaskNet6(nb_ref_imgs=4, output_exp=True), P['mask'], device), flow=_load(CM.Back2Future(nlevels=6), P['flow'], device))
    d = lambda t: t.to(device)
    errs, total = CE.flow_sample_errors(nets['disp'], nets['pose'], nets['mask'], nets['flow'], d(tgt), [d(r) for r in refs], d(K), d(Kinv),
                                        d(flow_gt), d(obj), THRESH=0.01)
    errs_o, total_o = OE.flow_sample_errors(P, tgt, refs, K, Kinv, flow_gt, obj, THRESH=0.01)
    assert np.allclose(np.array(errs, np.float64), np.array([float(e) for e in errs_o]), rtol=1e-3, atol=1e-4), (errs, errs_o)
    # the composed flow differs only where a census comparison |flow_cam - flow| < THRESH is a tie
    bad = ((total.cpu() - total_o).abs() > 1e-3 * total_o.abs().max()).float().mean().item()
    assert bad <= 1e-2, 'composed flow: %.2e of the pixels differ' % bad


EVAL_CASES = [case_eval_depth, case_eval_pose, case_eval_flow]
