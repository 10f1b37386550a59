"""Parity cases shared by the CPU-simulator tests (tests/test_sim_kernels.py) and the real-GPU tests
(tests/test_gpu_parity.py): each case runs the product path (cc_b200.* -> C ABI -> kernels) and the
oracle on the same seeded inputs on `device` and asserts agreement.

Tolerance: BASELINE.json's bar is 1e-4 relative in fp32, read as max-abs error relative to the
tensor's max-abs (tests/util.rel_err); masks / 0-1 targets must be bit-exact."""
import torch
from tests.util import golden, T, assert_close, assert_close_robust, rel_err
from cc_b200 import synth
from cc_b200 import loss_functions as CL, inverse_warp as CW, ssim as CS, pyramid as CP
from oracle import losses as OL, geometry as OG, ssim as OS

TOL = 1e-4


def dev_sample(B, H, W, seed, nlevels, device):
    s = synth.sample(B, H, W, seed=seed, nlevels=nlevels)
    out = {}
    for k, v in s.items():
        out[k] = [t.to(device) for t in v] if isinstance(v, list) else v.to(device)
    return out


def leafs(lst):
    return [x.detach().clone().requires_grad_(True) for x in lst]


def case_pyramid(device, B=2, H=64, W=96, nlevels=6):
    tgt, _ = synth.frames(B, H, W, seed=3)
    tgt = tgt.to(device)
    lv = CP.build(tgt, nlevels)
    for l in range(nlevels):
        ref = torch.nn.functional.adaptive_avg_pool2d(tgt, (H >> l, W >> l))
        assert_close(lv[l], ref, 2e-6, f'pyramid level {l}')


def case_warp_golden(device):
    """Stand-alone warp ops against the fixtures frozen from the reference (tests/golden)."""
    g = golden('warp_small')
    img, K, Kinv = T(g['img'], device), T(g['K'], device), T(g['Kinv'], device)
    wt = T(g['iw_wt'], device)
    mk = lambda k: T(g[k], device).clone().requires_grad_(True)
    for pm in ('zeros', 'border'):
        depth, pose = mk('depth'), mk('pose')
        out = CW.inverse_warp(img, depth, pose, K, Kinv, 'euler', pm)
        gd, gp = torch.autograd.grad((out * wt).sum(), [depth, pose])
        assert_close(out, g[f'iw_{pm}_out'], TOL, 'inverse_warp ' + pm)
        assert_close(gd, g[f'iw_{pm}_gdepth'], TOL, 'inverse_warp gdepth ' + pm)
        assert_close(gp, g[f'iw_{pm}_gpose'], TOL, 'inverse_warp gpose ' + pm)
        if pm == 'zeros':   # valid mask bit-exact
            assert torch.equal((out == 0).all(1), (T(g['iw_zeros_out'], device) == 0).all(1))
    depth, pose = mk('depth'), mk('pose')
    assert_close(CW.inverse_warp(img, depth, pose, K, Kinv, 'quat'), g['iw_quat_out'], TOL, 'quat')
    flow, imgl = mk('flow'), mk('img')
    fw = CW.flow_warp(imgl, flow)
    gf, gi = torch.autograd.grad((fw * wt).sum(), [flow, imgl])
    assert_close(fw, g['fw_out'], TOL, 'flow_warp')
    assert_close(gf, g['fw_gflow'], TOL, 'flow_warp gflow')
    assert_close(gi, g['fw_gimg'], TOL, 'flow_warp gimg')
    p2f = CW.pose2flow(depth, pose, K, Kinv)
    gd, gp = torch.autograd.grad((p2f * T(g['p2f_wt'], device)).sum(), [depth, pose])
    assert_close(p2f, g['p2f_out'], TOL, 'pose2flow')
    assert_close(gd, g['p2f_gdepth'], TOL, 'pose2flow gdepth')
    assert_close(gp, g['p2f_gpose'], TOL, 'pose2flow gpose')
    assert_close(CW.pose2flow(depth, pose, K, Kinv, padding_mode='zeros'), g['p2f_zeros_out'], TOL)
    assert_close(CW.pose_vec2mat(pose, 'euler'), g['posemat_euler'], 1e-6)
    assert_close(CW.pose_vec2mat(pose, 'quat'), g['posemat_quat'], 1e-6)
    assert torch.equal(CW.flow2oob(flow * 4).to(torch.uint8).cpu(), T(g['oob']))
    b_ = T(g['fw_out'], device).clone().requires_grad_(True)
    a_ = T(g['tgt'], device).clone().requires_grad_(True)
    sm = CS.ssim(a_, b_)
    assert_close(sm, g['ssim_out'], TOL, 'ssim map')
    g1, g2 = torch.autograd.grad((sm * T(g['ssim_wt'], device)).sum(), [a_, b_])
    assert_close(g2, g['ssim_gimg2'], TOL, 'ssim grad img2')
    a2, b2 = a_.detach().clone().requires_grad_(True), b_.detach().clone().requires_grad_(True)
    o1, o2 = torch.autograd.grad((OS.ssim(a2, b2) * T(g['ssim_wt'], device)).sum(), [a2, b2])
    assert_close(g1, o1, TOL, 'ssim grad img1')


def case_quat_backward(device):
    s = dev_sample(2, 24, 40, 11, 1, device)
    img, K, Kinv = s['refs'][0], s['K'], s['Kinv']
    wt = torch.randn(img.shape, generator=torch.Generator().manual_seed(5)).to(device)
    res = []
    for mod in (OG, CW):
        depth = s['depth'][0][:, 0].clone().requires_grad_(True)
        pose = (s['pose'][:, 0] * 3).clone().requires_grad_(True)
        out = mod.inverse_warp(img, depth, pose, K, Kinv, 'quat', 'zeros')
        res.append((out,) + torch.autograd.grad((out * wt).sum(), [depth, pose]))
    for a, b, nm in zip(res[1], res[0], ('out', 'gdepth', 'gpose')):
        assert_close(a, b, TOL, 'quat ' + nm)


def case_cfg0(device):
    """BASELINE.json configs[0]: inverse_warp + L1 photometric on one 3x128x416 triplet."""
    g = golden('cfg0')
    B, H, W = int(g['B']), int(g['H']), int(g['W'])
    tgt, refs = synth.frames(B, H, W, seed=20, n_refs=2)
    K, Kinv = synth.intrinsics(B, H, W)
    tgt, refs, K, Kinv = tgt.to(device), [r.to(device) for r in refs], K.to(device), Kinv.to(device)
    depth = synth.depths(B, H, W, 1, seed=21)[0][:, 0].to(device).requires_grad_(True)
    pose = synth.poses(B, 2, seed=22, big_tx_sample=False).to(device).requires_grad_(True)
    loss = 0
    for i in range(2):
        w = CW.inverse_warp(refs[i], depth, pose[:, i], K, Kinv)
        valid = 1 - (w == 0).prod(1, keepdim=True).type_as(w)
        loss = loss + ((tgt - w) * valid).abs().mean()
    gd, gp = torch.autograd.grad(loss, [depth, pose])
    assert_close(loss, g['loss'], TOL, 'cfg0 loss')
    assert_close(gd, g['gdepth'], TOL, 'cfg0 gdepth')
    assert_close(gp, g['gpose'], TOL, 'cfg0 gpose')


def _rigid(mod, s, wssim, lam, use_mask, NL, qch=0.5, pm='zeros'):
    depth, pose = leafs(s['depth'][:NL]), s['pose'].detach().clone().requires_grad_(True)
    em = leafs(s['emask'][:NL]) if use_mask else [None] * NL
    l = mod.photometric_reconstruction_loss(s['tgt'], s['refs'], s['K'], s['Kinv'], depth, em, pose,
                                            padding_mode=pm, wssim=wssim, lambda_oob=lam, qch=qch)
    return l, torch.autograd.grad(l, depth + [pose] + (em if use_mask else []))


def case_rigid_loss_golden(device):
    g = golden('loss_small')
    B, H, W, NL = int(g['B']), int(g['H']), int(g['W']), int(g['NL'])
    s = dev_sample(B, H, W, 30, NL, device)
    for tag, wssim in (('w997', 0.997), ('w0', 0.0)):
        for mtag in ('mask', 'nomask'):
            l, gr = _rigid(CL, s, wssim, 0.1 if tag == 'w0' else 0, mtag == 'mask', NL)
            key = f'rigid_{tag}_{mtag}'
            assert_close(l, g[key], TOL, key)
            for i in range(NL):
                assert_close(gr[i], g[f'{key}_gdepth{i}'], TOL, f'{key} gdepth{i}')
            assert_close(gr[NL], g[f'{key}_gpose'], TOL, key + ' gpose')
            if mtag == 'mask':
                for i in range(NL):
                    assert_close(gr[NL + 1 + i], g[f'{key}_gmask{i}'], TOL, f'{key} gmask{i}')


def case_rigid_loss_oracle(device, B=2, H=64, W=128, NL=4, seed=77, robust=False, oracle_device=None, product_first=False):
    s = dev_sample(B, H, W, seed, NL, device)
    so = s if oracle_device is None else dev_sample(B, H, W, seed, NL, oracle_device)
    chk = assert_close_robust if robust else assert_close
    combos = ((0.997, 0.0, True, 0.5, 'zeros'), (0.5, 0.2, False, 0.4, 'zeros'), (0.85, 0.0, True, 0.5, 'border'))
    pre = [_rigid(CL, s, *c[:3], NL, *c[3:]) for c in combos] if product_first else None
    for ci, (wssim, lam, use_mask, qch, pm) in enumerate(combos):
        lo, go = _rigid(OL, so, wssim, lam, use_mask, NL, qch, pm)
        lc, gc = pre[ci] if pre else _rigid(CL, s, wssim, lam, use_mask, NL, qch, pm)
        assert_close(lc, lo, TOL, f'rigid loss wssim={wssim}')
        for a, b in zip(gc, go):
            chk(a, b, TOL, what=f'rigid grad wssim={wssim} pm={pm}')


def _flow(mod, s, wssim, lam, use_mask, NL):
    ff, fb, em = leafs(s['flow_fwd'][:NL]), leafs(s['flow_bwd'][:NL]), leafs(s['emask'][:NL])
    fem = [1 - m[:, 1:3] for m in em] if use_mask else [None] * NL
    l = mod.photometric_flow_loss(s['tgt'], s['refs'][1:3], [fb, ff], fem, wssim=wssim, lambda_oob=lam)
    return l, torch.autograd.grad(l, ff + fb + (em if use_mask else []))


def case_flow_loss_golden(device):
    g = golden('loss_small')
    B, H, W, NL = int(g['B']), int(g['H']), int(g['W']), int(g['NL'])
    s = dev_sample(B, H, W, 30, NL, device)
    for tag, wssim in (('w997', 0.997), ('w0', 0.0)):
        l, gr = _flow(CL, s, wssim, 0, True, NL)
        key = f'flow_{tag}'
        assert_close(l, g[key], TOL, key)
        for i in range(NL):
            assert_close(gr[i], g[f'{key}_gff{i}'], TOL, f'{key} gff{i}')
            assert_close(gr[NL + i], g[f'{key}_gfb{i}'], TOL, f'{key} gfb{i}')
            assert_close(gr[2 * NL + i], g[f'{key}_gmask{i}'], TOL, f'{key} gmask{i}')
    s2 = dev_sample(2, 64, 128, 78, 4, device)
    lo, go = _flow(OL, s2, 0.6, 0.3, False, 4)
    lc, gc = _flow(CL, s2, 0.6, 0.3, False, 4)
    assert_close(lc, lo, TOL, 'flow loss nomask')
    for a, b in zip(gc, go):
        assert_close(a, b, TOL, 'flow grad nomask')


def case_occlusion_and_valid_masks(device, B=2, H=64, W=96, NL=3, seed=30):
    """The fused kernel's valid*(1-occ) map against the oracle's masks: bit-exact."""
    s = dev_sample(B, H, W, seed, NL, device)
    depth, pose = leafs(s['depth'][:NL]), s['pose'].detach().clone().requires_grad_(True)
    l = CL.photometric_reconstruction_loss(s['tgt'], s['refs'], s['K'], s['Kinv'], depth, [None] * NL, pose, wssim=0.0)
    vo = [t for t in l.grad_fn.keep if t.dim() == 4 and t.shape[1] == 4 and t.shape[0] == B][:0]
    d = l.grad_fn
    sizes = [(H >> i, W >> i) for i in range(NL)]
    vos = [t for t in d.keep if t.dim() == 4 and tuple(t.shape) in [(B, 4, h, w) for h, w in sizes]]
    assert len(vos) >= NL
    for lvl in range(NL):
        h, w = sizes[lvl]
        ds = s['tgt'].size(2) / h
        K_s = torch.cat((s['K'][:, 0:2] / ds, s['K'][:, 2:]), dim=1)
        Kinv_s = torch.cat((s['Kinv'][:, :, 0:2] * ds, s['Kinv'][:, :, 2:]), dim=2)
        occ = OL.depth_occlusion_masks(s['depth'][lvl], s['pose'], s['K'], s['Kinv'])
        refs_s = [torch.nn.functional.adaptive_avg_pool2d(r, (h, w)) for r in s['refs']]
        exp = []
        for i in range(4):
            wimg = OG.inverse_warp(refs_s[i], s['depth'][lvl][:, 0], s['pose'][:, i], K_s, Kinv_s)
            valid = 1 - (wimg == 0).prod(1).type_as(wimg)
            exp.append(valid * (1 - occ[:, i]))
        exp = torch.stack(exp, 1)
        got = [t for t in vos if tuple(t.shape) == (B, 4, h, w)][0]
        nm = (got != exp).sum().item()
        assert nm == 0, f'level {lvl}: {nm} valid/occlusion mask mismatches'


def case_smooth(device):
    g = golden('loss_small')
    B, H, W, NL = int(g['B']), int(g['H']), int(g['W']), int(g['NL'])
    s = dev_sample(B, H, W, 30, NL, device)
    for nm, preds in (('depth', s['depth']), ('flow', s['flow_fwd']), ('mask', s['emask'])):
        for fn, tag in ((lambda p: CL.edge_aware_smoothness_loss(s['tgt'], p), 'edge'), (CL.smooth_loss, 'smooth')):
            pl = leafs(preds)
            l = fn(pl)
            gr = torch.autograd.grad(l, pl)
            assert_close(l, g[f'{tag}_{nm}'], TOL, f'{tag}_{nm}')
            for i in range(NL):
                assert_close(gr[i], g[f'{tag}_{nm}_g{i}'], TOL, f'{tag}_{nm} grad{i}')


def case_bce_consensus(device):
    g = golden('loss_small')
    B, H, W, NL = int(g['B']), int(g['H']), int(g['W']), int(g['NL'])
    s = dev_sample(B, H, W, 30, NL, device)
    em = leafs(s['emask'])
    l = CL.explainability_loss(em)
    assert_close(l, g['expl'], TOL, 'explainability')
    for i, gg in enumerate(torch.autograd.grad(l, em)):
        assert_close(gg, g[f'expl_g{i}'], TOL, f'expl grad{i}')
    cam_f = [T(g[f'cam_f{i}'], device) for i in range(NL)]
    cam_b = [T(g[f'cam_b{i}'], device) for i in range(NL)]
    ff = [T(g[f'cons_ff{i}'], device) for i in range(NL)]
    fb = [T(g[f'cons_fb{i}'], device) for i in range(NL)]
    tg = CL.consensus_exp_masks(cam_f, cam_b, ff, fb, s['tgt'], s['refs'][2], s['refs'][1], wssim=0.997, wrig=1.0, ws=0.1)
    check_consensus_targets(tg, [T(g[f'cons_target{i}'], device) for i in range(NL)],
                            OL.consensus_sides(cam_f, cam_b, ff, fb, s['tgt'], s['refs'][2], s['refs'][1], wssim=0.997, wrig=1.0))
    tgr = [T(g[f'cons_target{i}'], device) for i in range(NL)]
    rig_f = [(a - b).abs() for a, b in zip(cam_f, ff)]
    rig_b = [(a - b).abs() for a, b in zip(cam_b, fb)]
    em = leafs(s['emask'])
    l = CL.consensus_depth_flow_mask(em, rig_b, rig_f, tgr, tgr, THRESH=0.5, wbce=0.3)
    assert_close(l, g['cdfm'], TOL, 'consensus_depth_flow_mask')
    for i, gg in enumerate(torch.autograd.grad(l, em)):
        assert_close(gg, g[f'cdfm_g{i}'], TOL, f'cdfm grad{i}')


def check_consensus_targets(got, ref, sides, what='consensus target', sides64=None):
    """0/1 targets must equal the reference's EXCEPT where the comparison `wrig*cam_err <= flow_err + 1e-8`
    (loss_functions.py:199-200) is a genuine tie: both sides are sums of 169-tap SSIM windows, so two correct fp32
    evaluations (separable vs 2-D window, CPU vs GPU) differ by a few 1e-7 relative and may land on either side.
    Every mismatching pixel has to be such a near-tie; there is no allowance for a fraction of arbitrary mismatches.
    The tie band is |lhs - rhs| <= 1e-5 * max(|lhs|, |rhs|, 1e-3), or - when `sides64` (the same formulas evaluated in
    fp64 on the same fp32 inputs) is given - the MEASURED noise of the reference's own fp32 evaluation: a pixel whose
    exact margin is below 4x the 99.99th percentile of |margin_fp32 - margin_fp64| of its level is decided by rounding
    in the reference itself (SSIM's sigma^2 = E[x^2] - mu^2 cancellation amplifies 1e-7 to ~1e-4 in flat regions).
    Returns (mismatches, pixels)."""
    n_mis = n_px = 0
    for i, (t, r, (lhs, rhs)) in enumerate(zip(got, ref, sides)):
        t, r, lhs, rhs = t.cpu(), r.cpu(), lhs.detach().cpu(), rhs.detach().cpu()
        assert set(torch.unique(t).tolist()) <= {0.0, 1.0}, f'{what} level {i}: not a 0/1 map'
        mis = t != r
        tie = (lhs - rhs).abs() <= 1e-5 * torch.maximum(torch.maximum(lhs.abs(), rhs.abs()), torch.tensor(1e-3))
        note = ''
        if sides64 is not None:
            l64, r64 = sides64[i][0].detach().cpu(), sides64[i][1].detach().cpu()
            m32, m64 = (lhs - rhs).double(), l64 - r64
            noise = float(torch.quantile((m32 - m64).abs().flatten()[:: max(1, m32.numel() // 1000000)], 0.9999))
            tie = tie | (m64.abs() <= 4 * noise)
            note = ' (fp32-vs-fp64 reference noise p99.99 %.2e)' % noise
            if int(mis.sum()):
                idx = mis.nonzero()[:8]
                print(f'{what} level {i}: {int(mis.sum())} flips{note}; exact margins of the first: ' +
                      ', '.join('%.2e' % float(m64[tuple(j)]) for j in idx))
        bad = int((mis & ~tie).sum())
        assert bad == 0, f'{what} level {i}: {bad} mismatching pixels that are not near-ties (of {int(mis.sum())} mismatches){note}'
        n_mis += int(mis.sum())
        n_px += t.numel()
    return n_mis, n_px


def case_consensus_fullsize(device, B=4, H=256, W=832, NL=6, seed=31):
    """consensus_exp_masks at the BASELINE size (b4 256x832, 6 levels) against the CPU oracle, tie-aware exact."""
    s = synth.sample(B, H, W, seed=seed, nlevels=NL)
    cam_f = [OG.pose2flow(d[:, 0], s['pose'][:, 2], s['K'], s['Kinv']) for d in s['depth']]
    cam_b = [OG.pose2flow(d[:, 0], s['pose'][:, 1], s['K'], s['Kinv']) for d in s['depth']]
    ff, fb = s['flow_fwd'], s['flow_bwd']
    dv = lambda x: [t.to(device) for t in x] if isinstance(x, list) else x.to(device)
    tg = CL.consensus_exp_masks(dv(cam_f), dv(cam_b), dv(ff), dv(fb), dv(s['tgt']), dv(s['refs'][2]), dv(s['refs'][1]),
                                wssim=0.997, wrig=1.0, ws=0.1)
    sides = OL.consensus_sides(cam_f, cam_b, ff, fb, s['tgt'], s['refs'][2], s['refs'][1], wssim=0.997, wrig=1.0)
    ref = [(lhs <= (rhs + 1e-8)).float() for lhs, rhs in sides]
    d64 = lambda x: [t.double() for t in x] if isinstance(x, list) else x.double()
    sides64 = OL.consensus_sides(d64(cam_f), d64(cam_b), d64(ff), d64(fb), d64(s['tgt']), d64(s['refs'][2]), d64(s['refs'][1]),
                                 wssim=0.997, wrig=1.0)
    n_mis, n_px = check_consensus_targets(tg, ref, sides, sides64=sides64)
    print('consensus targets b%d %dx%dx%d: %d near-tie flips of %d pixels' % (B, H, W, NL, n_mis, n_px))
    frac = sum(float(t.mean()) for t in tg) / NL
    assert 0.02 < frac < 0.98, 'degenerate consensus test (target fraction %.3f)' % frac


def case_loss_layer_fullsize(device, B=4, H=256, W=832, NL=6, seed=33, oracle_device=None):
    """Every fused loss kernel other than the rigid photometric one (covered by case_rigid_loss_oracle) at an
    arbitrary size against the oracle: photometric_flow_loss (mask + occlusion), edge-aware and second-order
    smoothness on 1/2/4-channel predictions, explainability BCE, consensus weighted BCE - loss and gradients."""
    od = device if oracle_device is None else oracle_device
    s, so = dev_sample(B, H, W, seed, NL, device), dev_sample(B, H, W, seed, NL, od)
    report = {}

    def both(name, fn_c, fn_o, leaves_c, leaves_o, robust=False):
        lc = fn_c(*leaves_c)
        gc = torch.autograd.grad(lc, [x for grp in leaves_c for x in grp])
        lo = fn_o(*leaves_o)
        go = torch.autograd.grad(lo, [x for grp in leaves_o for x in grp])
        assert_close(lc, lo, TOL, name + ' loss')
        worst = 0.0
        for a, b in zip(gc, go):
            (assert_close_robust if robust else assert_close)(a, b, TOL, what=name + ' grad')
            worst = max(worst, rel_err(a, b))
        report[name] = (rel_err(lc, lo), worst)

    for wssim in (0.997, 0.0):
        both('flow_loss_w%g' % wssim,
             lambda ff, fb, em: CL.photometric_flow_loss(s['tgt'], s['refs'][1:3], [fb, ff], [1 - m[:, 1:3] for m in em], wssim=wssim),
             lambda ff, fb, em: OL.photometric_flow_loss(so['tgt'], so['refs'][1:3], [fb, ff], [1 - m[:, 1:3] for m in em], wssim=wssim),
             [leafs(s['flow_fwd'][:NL]), leafs(s['flow_bwd'][:NL]), leafs(s['emask'][:NL])],
             [leafs(so['flow_fwd'][:NL]), leafs(so['flow_bwd'][:NL]), leafs(so['emask'][:NL])], robust=True)
    for nm in ('depth', 'flow_fwd', 'emask'):
        both('edge_' + nm, lambda p: CL.edge_aware_smoothness_loss(s['tgt'], p), lambda p: OL.edge_aware_smoothness_loss(so['tgt'], p),
             [leafs(s[nm][:NL])], [leafs(so[nm][:NL])])
        both('smooth_' + nm, lambda p: CL.smooth_loss(p), lambda p: OL.smooth_loss(p), [leafs(s[nm][:NL])], [leafs(so[nm][:NL])])
    both('explainability', lambda m: CL.explainability_loss(m), lambda m: OL.explainability_loss(m),
         [leafs(s['emask'][:NL])], [leafs(so['emask'][:NL])])
    tgt_c = [(torch.rand(B, 1, H >> l, W >> l, generator=torch.Generator().manual_seed(seed + l)) > 0.5).float() for l in range(NL)]
    rig = lambda d, key: [(a - b).abs() * 0.02 for a, b in zip(d[key][:NL], d['flow_bwd' if key == 'flow_fwd' else 'flow_fwd'][:NL])]
    both('consensus_bce',
         lambda m: CL.consensus_depth_flow_mask(m, rig(s, 'flow_bwd'), rig(s, 'flow_fwd'), [t.to(device) for t in tgt_c], [t.to(device) for t in tgt_c], THRESH=0.01, wbce=0.5),
         lambda m: OL.consensus_depth_flow_mask(m, rig(so, 'flow_bwd'), rig(so, 'flow_fwd'), [t.to(od) for t in tgt_c], [t.to(od) for t in tgt_c], THRESH=0.01, wbce=0.5),
         [leafs(s['emask'][:NL])], [leafs(so['emask'][:NL])])
    print('loss layer b%d %dx%dx%d rel err (loss, worst grad): ' % (B, H, W, NL) + ', '.join('%s %.1e/%.1e' % (k, a, b) for k, (a, b) in report.items()))


def case_asserts(device):
    import pytest
    z = torch.zeros
    with pytest.raises(AssertionError, match='wrong size for depth, expected BxHxW'):
        CW.inverse_warp(z(1, 3, 4, 4, device=device), z(1, 1, 4, 4, device=device), z(1, 6, device=device),
                        torch.eye(3, device=device)[None], torch.eye(3, device=device)[None])
    with pytest.raises(AssertionError, match='wrong size for flow, expected Bx2xHxW'):
        CW.flow_warp(z(1, 3, 4, 4, device=device), z(1, 3, 4, 4, device=device))


ALL_CASES = [case_pyramid, case_warp_golden, case_quat_backward, case_cfg0, case_rigid_loss_golden,
             case_rigid_loss_oracle, case_flow_loss_golden, case_occlusion_and_valid_masks, case_smooth,
             case_bce_consensus, case_asserts]
