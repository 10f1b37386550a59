"""The oracle (CPU restatement) against fixtures frozen from the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import torch
import pytest
from tests.util import golden, T, assert_close, key_with_stride, pick
from cc_b200 import synth
from oracle import geometry as G, ssim as S, losses as L, nets as N, step as ST

TOL = 2e-6   # same torch ops in the same order: expect ~bit-equality


def leaf(a):
    return T(a).clone().requires_grad_(True)


def test_warp_layer():
    g = golden('warp_small')
    img, K, Kinv = T(g['img']), T(g['K']), T(g['Kinv'])
    depth, pose, flow = leaf(g['depth']), leaf(g['pose']), leaf(g['flow'])
    wt = T(g['iw_wt'])
    for pm in ('zeros', 'border'):
        out = G.inverse_warp(img, depth, pose, K, Kinv, 'euler', pm)
        gd, gp = torch.autograd.grad((out * wt).sum(), [depth, pose])
        assert_close(out, g[f'iw_{pm}_out'], TOL, pm)
        assert_close(gd, g[f'iw_{pm}_gdepth'], TOL)
        assert_close(gp, g[f'iw_{pm}_gpose'], TOL)
    assert_close(G.inverse_warp(img, depth, pose, K, Kinv, 'quat'), g['iw_quat_out'], TOL)
    imgl = leaf(g['img'])
    fw = G.flow_warp(imgl, flow)
    gf, gi = torch.autograd.grad((fw * wt).sum(), [flow, imgl])
    assert_close(fw, g['fw_out'], TOL)
    assert_close(gf, g['fw_gflow'], TOL)
    assert_close(gi, g['fw_gimg'], TOL)
    p2f = G.pose2flow(depth, pose, K, Kinv)
    gd, gp = torch.autograd.grad((p2f * T(g['p2f_wt'])).sum(), [depth, pose])
    assert_close(p2f, g['p2f_out'], TOL)
    assert_close(gd, g['p2f_gdepth'], TOL)
    assert_close(gp, g['p2f_gpose'], TOL)
    assert_close(G.pose2flow(depth, pose, K, Kinv, padding_mode='zeros'), g['p2f_zeros_out'], TOL)
    assert_close(G.pose_vec2mat(pose, 'euler'), g['posemat_euler'], TOL)
    assert_close(G.pose_vec2mat(pose, 'quat'), g['posemat_quat'], TOL)
    assert torch.equal(G.flow2oob(flow * 4).to(torch.uint8), T(g['oob']))
    b_ = leaf(g['fw_out'])
    sm = S.ssim(T(g['tgt']), b_)
    assert_close(sm, g['ssim_out'], TOL)
    assert_close(torch.autograd.grad((sm * T(g['ssim_wt'])).sum(), [b_])[0], g['ssim_gimg2'], TOL)


def test_size_assert_message():
    with pytest.raises(AssertionError, match='wrong size for depth, expected BxHxW'):
        G.inverse_warp(torch.zeros(1, 3, 4, 4), torch.zeros(1, 1, 4, 4), torch.zeros(1, 6),
                       torch.eye(3)[None], torch.eye(3)[None])


def test_cfg0():
    g = golden('cfg0')
    B, H, W = int(g['B']), int(g['H']), int(g['W'])
    tgt, refs = synth.frames(B, H, W, seed=20, n_refs=2)
    K, Kinv = synth.intrinsics(B, H, W)
    depth = synth.depths(B, H, W, 1, seed=21)[0][:, 0].clone().requires_grad_(True)
    pose = synth.poses(B, 2, seed=22, big_tx_sample=False).clone().requires_grad_(True)
    loss = 0
    for i in range(2):
        w = G.inverse_warp(refs[i], depth, pose[:, i], K, Kinv)
        valid = 1 - (w == 0).prod(1, keepdim=True).type_as(w)
        loss = loss + ((tgt - w) * valid).abs().mean()
    gd, gp = torch.autograd.grad(loss, [depth, pose])
    assert_close(loss, g['loss'], TOL)
    assert_close(gd, g['gdepth'], TOL)
    assert_close(gp, g['gpose'], TOL)


def test_loss_layer():
    g = golden('loss_small')
    B, H, W, NL = int(g['B']), int(g['H']), int(g['W']), int(g['NL'])
    s = synth.sample(B, H, W, seed=30, nlevels=NL)
    tgt, refs, K, Kinv = s['tgt'], s['refs'], s['K'], s['Kinv']
    lf = lambda lst: [x.clone().requires_grad_(True) for x in lst]
    for tag, wssim in (('w997', 0.997), ('w0', 0.0)):
        for mtag in ('mask', 'nomask'):
            depth, pose = lf(s['depth']), s['pose'].clone().requires_grad_(True)
            em = lf(s['emask']) if mtag == 'mask' else [None] * NL
            l = L.photometric_reconstruction_loss(tgt, refs, K, Kinv, depth, em, pose, wssim=wssim,
                                                  lambda_oob=0.1 if tag == 'w0' else 0)
            gr = torch.autograd.grad(l, depth + [pose] + (em if mtag == 'mask' else []))
            key = f'rigid_{tag}_{mtag}'
            assert_close(l, g[key], TOL, key)
            for i in range(NL):
                assert_close(gr[i], g[f'{key}_gdepth{i}'], TOL, key)
            assert_close(gr[NL], g[f'{key}_gpose'], TOL, key)
            if mtag == 'mask':
                for i in range(NL):
                    assert_close(gr[NL + 1 + i], g[f'{key}_gmask{i}'], TOL, key)
        ff, fb, em = lf(s['flow_fwd']), lf(s['flow_bwd']), lf(s['emask'])
        l = L.photometric_flow_loss(tgt, refs[1:3], [fb, ff], [1 - m[:, 1:3] for m in em], wssim=wssim)
        gr = torch.autograd.grad(l, ff + fb + em)
        key = f'flow_{tag}'
        assert_close(l, g[key], TOL, key)
        for i in range(NL):
            assert_close(gr[i], g[f'{key}_gff{i}'], TOL)
            assert_close(gr[NL + i], g[f'{key}_gfb{i}'], TOL)
            assert_close(gr[2 * NL + i], g[f'{key}_gmask{i}'], TOL)
    for lv in (0, 1):
        assert torch.equal(L.depth_occlusion_masks(s['depth'][lv], s['pose'], K, Kinv), T(g[f'depth_occ{lv}']))
    for nm, preds in (('depth', s['depth']), ('flow', s['flow_fwd']), ('mask', s['emask'])):
        for fn, tag in ((lambda p: L.edge_aware_smoothness_loss(tgt, p), 'edge'), (L.smooth_loss, 'smooth')):
            pl = lf(preds)
            l = fn(pl)
            gr = torch.autograd.grad(l, pl)
            assert_close(l, g[f'{tag}_{nm}'], TOL)
            for i in range(NL):
                assert_close(gr[i], g[f'{tag}_{nm}_g{i}'], TOL)
    em = lf(s['emask'])
    l = L.explainability_loss(em)
    assert_close(l, g['expl'], TOL)
    for i, gg in enumerate(torch.autograd.grad(l, em)):
        assert_close(gg, g[f'expl_g{i}'], TOL)
    cam_f = [T(g[f'cam_f{i}']) for i in range(NL)]
    cam_b = [T(g[f'cam_b{i}']) for i in range(NL)]
    for i in range(NL):
        assert_close(G.pose2flow(s['depth'][i].squeeze(1), s['pose'][:, 2], K, Kinv), cam_f[i], TOL)
    ff = [T(g[f'cons_ff{i}']) for i in range(NL)]
    fb = [T(g[f'cons_fb{i}']) for i in range(NL)]
    tg = L.consensus_exp_masks(cam_f, cam_b, ff, fb, tgt, refs[2], refs[1], wssim=0.997, wrig=1.0, ws=0.1)
    for i in range(NL):
        assert torch.equal(tg[i], T(g[f'cons_target{i}']))
        assert 0.02 < tg[i].mean().item() < 0.98      # both classes present
    rig_f = [(a - b).abs() for a, b in zip(cam_f, ff)]
    rig_b = [(a - b).abs() for a, b in zip(cam_b, fb)]
    em = lf(s['emask'])
    l = L.consensus_depth_flow_mask(em, rig_b, rig_f, tg, tg, THRESH=0.5, wbce=0.3)
    assert_close(l, g['cdfm'], TOL)
    for i, gg in enumerate(torch.autograd.grad(l, em)):
        assert_close(gg, g[f'cdfm_g{i}'], TOL)


def _check_grads(g, prefix, names, params, grads, tol):
    for n, gg in zip(names, grads):
        k, stride = key_with_stride(g, prefix + n)
        assert_close(pick(gg, stride), g[k], tol, k)


NTOL = 2e-5   # conv algorithm choice (mkldnn) may differ between module and functional calls


def test_nets():
    g = golden('nets_small')
    B, H, W = 2, 64, 128
    tgt, refs = synth.frames(B, H, W, seed=40)
    wts = lambda shape, seed: torch.randn(shape, generator=torch.Generator().manual_seed(seed))
    P = N.clone_params(N.disp_params(), requires_grad=True)
    disps = N.disp_forward(P, tgt, training=True)
    for i, x in enumerate(disps):
        assert_close(x, g[f'disp_out{i}'], NTOL, f'disp{i}')
    names = ['conv1.0.weight', 'conv1.2.bias', 'conv2.0.conv1.weight', 'conv2.0.downsample.0.weight',
             'conv2.0.downsample.1.weight', 'conv2.0.downsample.1.bias', 'conv7.1.conv2.weight',
             'upconv7.0.weight', 'upconv1.0.bias', 'iconv1.0.conv1.weight', 'iconv3.0.downsample.0.weight',
             'predict_disp1.0.weight', 'predict_disp6.0.bias']
    loss = sum((x * wts(x.shape, 50 + i)).sum() for i, x in enumerate(disps))
    _check_grads(g, 'disp_g_', names, P, torch.autograd.grad(loss, [P[n] for n in names]), 2e-4)
    assert_close(P['conv2.0.downsample.1.running_mean'], g['disp_rm'], NTOL)
    assert_close(P['iconv1.0.downsample.1.running_var'], g['disp_rv'], NTOL)
    with torch.no_grad():
        assert_close(N.disp_forward(P, tgt, training=False), g['disp_eval'], NTOL)
        t2, _ = synth.frames(2, 24, 40, seed=41)
        P2 = N.clone_params(N.disp_params())
        for i, x in enumerate(N.disp_forward(P2, t2, training=True)):
            assert_close(x, g[f'disp_odd_out{i}'], NTOL)
    Pp = N.clone_params(N.pose_params(), requires_grad=True)
    pose = N.pose_forward(Pp, tgt, refs)
    assert_close(pose, g['pose_out'], NTOL)
    pn = ['conv1.0.weight', 'conv2.0.weight', 'conv8.0.bias', 'pose_pred.weight', 'pose_pred.bias']
    _check_grads(g, 'pose_g_', pn, Pp, torch.autograd.grad((pose * wts(pose.shape, 60)).sum(), [Pp[n] for n in pn]), 2e-4)
    tgt, refs = synth.frames(1, 64, 64, seed=42)
    Pm = N.clone_params(N.mask_params(), requires_grad=True)
    ms = N.mask_forward(Pm, tgt, refs)
    for i, x in enumerate(ms):
        assert_close(x, g[f'mask_out{i}'], NTOL)
    mn = ['conv1.0.weight', 'conv6.0.weight', 'deconv6.0.weight', 'deconv1.0.weight', 'deconv3.0.bias',
          'pred_mask1.weight', 'pred_mask6.bias']
    _check_grads(g, 'mask_g_', mn, Pm,
                 torch.autograd.grad(sum((x * wts(x.shape, 70 + i)).sum() for i, x in enumerate(ms)), [Pm[n] for n in mn]), 2e-4)
    Pf = N.clone_params(N.flow_params(), requires_grad=True)
    ff, fb, occ = N.flow_forward(Pf, tgt, refs[1:3], training=True)
    for i in range(6):
        assert_close(ff[i], g[f'flow_fwd{i}'], 1e-4, f'ff{i}')
        assert_close(fb[i], g[f'flow_bwd{i}'], 1e-4, f'fb{i}')
    assert_close(occ[0][:, :, ::4, ::4], g['flow_occ0'], 1e-4)
    assert_close(occ[5], g['flow_occ5'], 1e-4)
    fn = ['conv1a.0.weight', 'conv1b.2.bias', 'conv6c.0.weight', 'decoder_fwd6.0.weight',
          'decoder_bwd2.10.weight', 'decoder_fwd2.0.weight', 'decoder_bwd4.4.bias']
    lossf = sum((x * wts(x.shape, 80 + i)).sum() + (y * wts(x.shape, 80 + i)).sum() * 0.5
                for i, (x, y) in enumerate(zip(ff, fb)))
    _check_grads(g, 'flow_g_', fn, Pf, torch.autograd.grad(lossf, [Pf[n] for n in fn]), 5e-4)
    with torch.no_grad():
        e = N.flow_forward(Pf, tgt, refs[1:3], training=False)
        assert_close(e[0][:, :, ::2, ::2], g['flow_eval_fwd'], 1e-4)


def test_step_body():
    g = golden('step_small')
    B, H, W = 2, 64, 128
    tgt, refs = synth.frames(B, H, W, seed=40)
    K, Kinv = synth.intrinsics(B, H, W)
    P = ST.make_params('cfg3')
    loss, aux = ST.loss_cfg3(P, tgt, refs, K, Kinv)
    loss.backward()
    for k in ('loss_1', 'loss_2', 'loss_3', 'loss_4', 'loss_5'):
        assert_close(aux[k], g[k], 1e-4, k)
    assert_close(loss, g['loss'], 1e-4)
    for nm in ('disp', 'pose', 'mask', 'flow'):
        gn = torch.sqrt(sum((t.grad ** 2).sum() for t in P[nm].values() if t.requires_grad and t.grad is not None))
        assert_close(gn, g[f'gnorm_{nm}'], 2e-3, 'gnorm_' + nm)
    k, st = key_with_stride(g, 'g_disp_conv1.0.weight')
    assert_close(pick(P['disp']['conv1.0.weight'].grad, st), g[k], 2e-3)
    k, st = key_with_stride(g, 'g_pose_pose_pred.bias')
    assert_close(pick(P['pose']['pose_pred.bias'].grad, st), g[k], 2e-3)
    P1 = ST.make_params('cfg1')
    loss, aux = ST.loss_cfg1(P1, tgt, refs, K, Kinv)
    loss.backward()
    assert_close(loss, g['cfg1_loss'], 1e-4)
    assert_close(aux['loss_1'], g['cfg1_l1'], 1e-4)
    assert_close(P1['disp']['conv1.0.weight'].grad, g['cfg1_g_disp_conv1.0.weight'], 2e-3)
    assert_close(P1['pose']['pose_pred.bias'].grad, g['cfg1_g_pose_pose_pred.bias'], 2e-3)


def test_oracle_adam_matches_torch():
    torch.manual_seed(0)
    w = torch.randn(7, 5)
    a = w.clone().requires_grad_(True)
    b = w.clone().requires_grad_(True)
    oa, ob = ST.Adam([a], 1e-2), torch.optim.Adam([b], 1e-2, betas=(0.9, 0.999))
    for _ in range(5):
        for p, o in ((a, oa), (b, ob)):
            o.zero_grad()
            (p.sin() * p).sum().backward()
            o.step()
    assert_close(a, b, 1e-6)


def test_validation_metrics():
    """oracle/metrics.py against the reference's loss_functions.py:355-467 (fixture frozen from the reference)."""
    from oracle import metrics as OM
    g = golden('metrics_small')
    gt, pr, pn, mask = T(g['gt']), T(g['pr']), T(g['pn']), T(g['mask'])
    assert_close(OM.flow_diff(gt, pr), g['flow_diff'], 1e-6)
    assert abs(OM.compute_epe(gt, pr) - float(g['epe3'])) <= 1e-6 * float(g['epe3'])
    assert abs(OM.compute_epe(gt[:, :2].contiguous(), pr) - float(g['epe2'])) <= 1e-6 * float(g['epe2'])
    assert abs(OM.outlier_err(gt, pr * 3) - float(g['outlier'])) <= 1e-6
    assert_close(torch.tensor(OM.compute_all_epes(gt, pr, pn, mask)), g['all_epes'], 1e-6)
    assert_close(torch.tensor(OM.compute_all_epes(gt, pr, pn, mask, THRESH=0.3)), g['all_epes_t3'], 1e-6)
    for crop, key in ((True, 'errors_crop'), (False, 'errors_nocrop')):
        assert_close(torch.stack([torch.as_tensor(v) for v in OM.compute_errors(T(g['dgt']), T(g['dpr']), crop=crop)]), g[key], 1e-6)


def _transform_params(g, key_random, seed_np, B, Hs, Ws, scale_crop):
    import random
    import numpy as np
    from cc_b200.input_pipeline import draw_params
    # the reference draws flip and scale/crop decisions sample by sample in Compose order: same generators, same order
    random.seed(int(g[key_random]))
    if seed_np is not None:
        np.random.seed(int(seed_np))
    return draw_params(B, Hs, Ws, scale_crop=scale_crop)


def test_input_transforms():
    """oracle/transforms.py (and the host-side parameter draw / intrinsics update of cc_b200.input_pipeline) against the
    reference's Compose([RandomHorizontalFlip, RandomScaleCrop, ArrayToTensor, Normalize]) run on uint8 frames."""
    import numpy as np
    from oracle import transforms as OT
    from cc_b200.input_pipeline import augment_intrinsics
    g = golden('transforms_small')
    frames, K = g['frames'], g['K']
    B, F, Hs, Ws, _ = frames.shape
    p = _transform_params(g, 'seed_random', g['seed_np'], B, Hs, Ws, True)
    out, Kb = OT.apply(frames, K, p)
    assert np.array_equal(Kb, g['K_out']), 'intrinsics after flip + scale-crop must match the reference bit for bit'
    assert np.array_equal(augment_intrinsics(np.broadcast_to(K, (B, 3, 3)), p, Ws), g['K_out'])
    assert p['flip'].sum() > 0 and (p['scaled_w'] > Ws).any()
    # PIL rounds the resampled image to uint8 after each of its two passes: <= 1 step per pass, in (v/255-.5)/.5 units
    assert np.abs(out - g['out']).max() <= 2 * (2 / 255) + 1e-6
    assert np.abs(out - g['out']).mean() <= 0.6 * (2 / 255)
    p2 = _transform_params(g, 'seed_flip', None, B, Hs, Ws, False)
    out2, K2 = OT.apply(frames, K, p2)
    assert np.array_equal(out2, g['out_flip']) and np.array_equal(K2, g['K_flip'])
