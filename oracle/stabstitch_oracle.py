"""CPU oracle for the StabStitch++ per-frame inference hot path.

TEST INFRASTRUCTURE ONLY.  This file is a CPU restatement (torch-CPU fp32 functional ops +
an fp64 arbiter for the thin-plate-spline) of the reference algorithm.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
it, and only as the checker / the timed CPU baseline - the product path
(`stabstitch2_b200`) never imports it and fails loudly when its CUDA library is missing.

Parity pinning: the reference (nie-lang/StabStitch2 @ 646dad0) ships NO tests, golden
vectors or fixtures for this path, so the oracle is pinned against
OUTPUTS OF THE REFERENCE ITSELF, imported unmodified in the build container
(tests/golden/ref_harness.py): tests/golden/make_golden.py runs the reference on seeded inputs
(container only, /root/reference does not travel) and commits its outputs as
tests/golden/*.npz; tests/test_oracle_golden.py checks every oracle function against them.

All file:line citations are relative to /root/reference/Full_model_inference/Codes/.
Weights travel as a plain `state_dict` (name -> tensor) with the reference's key names.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

GRID_H = 6  # grid_res.py:3
GRID_W = 8  # grid_res.py:4
NPT = (GRID_H + 1) * (GRID_W + 1)  # 63 control points


# --------------------------------------------------------------------------------------
# mesh helpers
# --------------------------------------------------------------------------------------
def rigid_mesh(bs, height, width):
    """[bs,7,9,2] (x,y) vertices of the undeformed grid.  spatial_network.py:39-50,
    test_online_tra.py:71-83."""
    xs = torch.linspace(0.0, float(width), GRID_W + 1)
    ys = torch.linspace(0.0, float(height), GRID_H + 1)
    m = torch.empty(GRID_H + 1, GRID_W + 1, 2)
    m[..., 0] = xs[None, :]
    m[..., 1] = ys[:, None]
    return m[None].expand(bs, -1, -1, -1)


def norm_mesh(mesh, height, width):
    """pixel -> [-1,1].  spatial_network.py:53-59, test_online_tra.py:85-91.
    `height`/`width` may be python floats or 0-dim tensors (float canvas extent)."""
    bs = mesh.shape[0]
    nx = mesh[..., 0] * 2.0 / float(width) - 1.0
    ny = mesh[..., 1] * 2.0 / float(height) - 1.0
    return torch.stack([nx, ny], -1).reshape(bs, -1, 2)


def recover_mesh(nmesh, height, width):
    """[-1,1] -> pixel, [bs,63,2] -> [bs,7,9,2].  test_online_tra.py:61-69."""
    bs = nmesh.shape[0]
    x = (nmesh[..., 0] + 1) * float(width) / 2.0
    y = (nmesh[..., 1] + 1) * float(height) / 2.0
    return torch.stack([x, y], 2).reshape(bs, GRID_H + 1, GRID_W + 1, 2)


# --------------------------------------------------------------------------------------
# 4-point homography (utils/torch_DLT.py:17-45) and the bidirectional split
# --------------------------------------------------------------------------------------
def tensor_dlt(src_p, dst_p):
    """Solve the 8x8 DLT system with an explicit fp32 inverse; H maps src -> dst."""
    bs = src_p.shape[0]
    A = torch.zeros(bs, 8, 8)
    b = torch.zeros(bs, 8, 1)
    for p in range(4):
        x, y = src_p[:, p, 0], src_p[:, p, 1]
        u, v = dst_p[:, p, 0], dst_p[:, p, 1]
        A[:, 2 * p, 0], A[:, 2 * p, 1], A[:, 2 * p, 2] = x, y, 1.0
        A[:, 2 * p, 6], A[:, 2 * p, 7] = -(u * x), -(u * y)
        A[:, 2 * p + 1, 3], A[:, 2 * p + 1, 4], A[:, 2 * p + 1, 5] = x, y, 1.0
        A[:, 2 * p + 1, 6], A[:, 2 * p + 1, 7] = -(v * x), -(v * y)
        b[:, 2 * p, 0], b[:, 2 * p + 1, 0] = u, v
    h8 = torch.matmul(torch.inverse(A), b).reshape(bs, 8)
    return torch.cat([h8, torch.ones(bs, 1)], 1).reshape(bs, 3, 3)


def h2mesh(H, rigid):
    """Apply H^-1 to the 63 rigid vertices and de-homogenise.  spatial_network.py:20-36."""
    bs = rigid.shape[0]
    Hinv = torch.inverse(H)
    pts = torch.cat([rigid.reshape(bs, -1, 2), torch.ones(bs, NPT, 1)], 2)
    q = torch.matmul(Hinv, pts.permute(0, 2, 1))
    x = q[:, 0, :] / q[:, 2, :]
    y = q[:, 1, :] / q[:, 2, :]
    return torch.stack([x, y], 2).reshape(bs, GRID_H + 1, GRID_W + 1, 2)


def homography_split(h_motion, img_h, img_w, scale=1.0):
    """H, H_tgt (half motion), H_ref = H^-1 H_tgt.  spatial_network.py:73-93 (scale 1) and
    :291-300 (scale 1/8 -> pass scale=8)."""
    bs = h_motion.shape[0]
    src = torch.tensor([[0.0, 0.0], [img_w, 0.0], [0.0, img_h], [img_w, img_h]])
    src = src[None].expand(bs, -1, -1)
    dst = src + h_motion
    dst_tgt = src + h_motion / 2.0
    H = tensor_dlt(src / scale, dst / scale) if scale != 1.0 else tensor_dlt(src, dst)
    H_tgt = tensor_dlt(src / scale, dst_tgt / scale) if scale != 1.0 else tensor_dlt(src, dst_tgt)
    H_ref = torch.matmul(torch.inverse(H), H_tgt)
    return H, H_ref, H_tgt


# --------------------------------------------------------------------------------------
# the 4-tap sampler shared by both transformers
# (utils/torch_tps_transform.py:30-106 == utils/torch_homo_transform.py:50-125)
# --------------------------------------------------------------------------------------
def bilinear_sample(im, xs, ys):
    """im [bn,C,H,W]; xs, ys [bn, P] normalised source coords -> [bn, C, P].

    x=(xs+1)*W/2 (NOT align-corners); x0=floor, x1=x0+1, both clamped; weights from the
    CLAMPED integer coords; out = ((wa*Ia + wb*Ib) + wc*Ic) + wd*Id with
    a=(y0,x0) b=(y1,x0) c=(y0,x1) d=(y1,x1).  Out-of-image samples therefore cancel to
    rounding residue, not to exact zero - restated faithfully here."""
    bn, C, H, W = im.shape
    x = (xs + 1.0) * W / 2.0
    y = (ys + 1.0) * H / 2.0
    x0 = torch.floor(x).to(torch.int64)
    y0 = torch.floor(y).to(torch.int64)
    x1 = (x0 + 1).clamp(0, W - 1)
    y1 = (y0 + 1).clamp(0, H - 1)
    x0 = x0.clamp(0, W - 1)
    y0 = y0.clamp(0, H - 1)
    flat = im.reshape(bn, C, H * W)

    def tap(yy, xx):
        idx = (yy * W + xx)[:, None, :].expand(-1, C, -1)
        return torch.gather(flat, 2, idx)

    Ia, Ib, Ic, Id = tap(y0, x0), tap(y1, x0), tap(y0, x1), tap(y1, x1)
    x0f, x1f, y0f, y1f = x0.float(), x1.float(), y0.float(), y1.float()
    wa = ((x1f - x) * (y1f - y))[:, None, :]
    wb = ((x1f - x) * (y - y0f))[:, None, :]
    wc = ((x - x0f) * (y1f - y))[:, None, :]
    wd = ((x - x0f) * (y - y0f))[:, None, :]
    return wa * Ia + wb * Ib + wc * Ic + wd * Id


def homo_warp(U, theta, out_size):
    """Backward homography warp.  utils/torch_homo_transform.py:128-180."""
    bn, C, H, W = U.shape
    Ho, Wo = out_size
    xt = torch.linspace(-1.0, 1.0, Wo)[None, :].expand(Ho, Wo).reshape(-1)
    yt = torch.linspace(-1.0, 1.0, Ho)[:, None].expand(Ho, Wo).reshape(-1)
    grid = torch.stack([xt, yt, torch.ones_like(xt)], 0)  # [3, P]
    tg = torch.matmul(theta.reshape(-1, 3, 3).float(), grid[None])  # [bn,3,P]
    xs, ys, ts = tg[:, 0], tg[:, 1], tg[:, 2]
    ts = ts + 1e-6 * (1.0 - (ts.abs() >= 1e-7).float())  # :167-170
    out = bilinear_sample(U, xs / ts, ys / ts)
    return out.reshape(bn, C, Ho, Wo)


# --------------------------------------------------------------------------------------
# thin-plate spline (utils/torch_tps_transform.py, utils/torch_tps_transform_point.py)
# --------------------------------------------------------------------------------------
def tps_system_matrix(source):
    """fp32 [bn,66,66] system matrix W = [[P, K],[0, P^T]].  torch_tps_transform.py:168-198."""
    bn, n, _ = source.shape
    p = torch.cat([torch.ones(bn, n, 1), source], 2)
    diff = p[:, :, None, :] - p[:, None, :, :]
    d2 = (diff * diff).sum(3)
    K = d2 * torch.log(d2 + 1e-6)
    top = torch.cat([p, K], 2)
    bot = torch.cat([torch.zeros(bn, 3, 3), p.permute(0, 2, 1)], 2)
    return torch.cat([top, bot], 1)


def tps_solve(source, target):
    """T [bn,2,66] fp32: coefficient order [a0, ax, ay, w_0..w_62].  The system is inverted
    EXPLICITLY in fp64 and multiplied (torch_tps_transform.py:206-226)."""
    bn = source.shape[0]
    W = tps_system_matrix(source)
    Winv = torch.inverse(W.double())
    tp = torch.cat([target, torch.zeros(bn, 3, 2)], 1).double()
    return torch.matmul(Winv, tp).permute(0, 2, 1).float()


def tps_eval(T, source, xt, yt):
    """Evaluate the spline at points (xt, yt) [P] or [bn,P] -> xs, ys [bn,P].
    torch_tps_transform.py:108-149 (grid rows [1, x, y, r_0..r_62], T x grid)."""
    bn = source.shape[0]
    if xt.dim() == 1:
        xt = xt[None].expand(bn, -1)
        yt = yt[None].expand(bn, -1)
    px = source[:, :, 0:1]
    py = source[:, :, 1:2]
    d2 = torch.square(xt[:, None, :] - px) + torch.square(yt[:, None, :] - py)
    r = d2 * torch.log(d2 + 1e-6)
    grid = torch.cat([torch.ones(bn, 1, xt.shape[1]), xt[:, None, :], yt[:, None, :], r], 1)
    tg = torch.matmul(T, grid)
    return tg[:, 0], tg[:, 1]


def tps_warp(U, source, target, out_size, mode="NORMAL"):
    """utils/torch_tps_transform.py:7 `transformer`.  U [bn,C,H,W] -> [bn,C,Ho,Wo]."""
    bn, C, H, W = U.shape
    Ho, Wo = int(out_size[0]), int(out_size[1])
    T = tps_solve(source, target)
    xt = torch.linspace(-1.0, 1.0, Wo)[None, :].expand(Ho, Wo).reshape(-1)
    yt = torch.linspace(-1.0, 1.0, Ho)[:, None].expand(Ho, Wo).reshape(-1)
    xs, ys = tps_eval(T, source, xt, yt)
    if mode == "NORMAL":
        return bilinear_sample(U.float(), xs, ys).reshape(bn, C, Ho, Wo)
    g = torch.stack([xs.reshape(bn, Ho, Wo), ys.reshape(bn, Ho, Wo)], 3)
    return F.grid_sample(U, g, align_corners=True)  # :158-162 (FAST)


def tps_point(point, source, target):
    """utils/torch_tps_transform_point.py:6 `transformer`: move `point` [bn,63,2]."""
    T = tps_solve(source, target)
    xs, ys = tps_eval(T, source, point[:, :, 0], point[:, :, 1])
    return torch.stack([xs, ys], 2)


def tps_source_coords_fp64(source, target, Ho, Wo, W_in, H_in):
    """fp64 ARBITER: source-pixel coordinates (x, y) [bn,Ho,Wo] of every canvas pixel, all
    arithmetic in float64 from the fp32 inputs (the pixel lattice is the exact
    -1 + 2 i/(n-1)).  Used to decide who is closer to the truth when two fp32 evaluations
    disagree (SURVEY.md Hard part 2)."""
    s = source.double().numpy()
    t = target.double().numpy()
    bn, n, _ = s.shape
    out_x = np.empty((bn, Ho, Wo))
    out_y = np.empty((bn, Ho, Wo))
    xt = -1.0 + 2.0 * np.arange(Wo) / max(Wo - 1, 1)
    yt = -1.0 + 2.0 * np.arange(Ho) / max(Ho - 1, 1)
    for b in range(bn):
        p = np.concatenate([np.ones((n, 1)), s[b]], 1)
        d2 = ((p[:, None, :] - p[None, :, :]) ** 2).sum(2)
        K = d2 * np.log(d2 + 1e-6)
        Wm = np.zeros((n + 3, n + 3))
        Wm[:n, :3] = p
        Wm[:n, 3:] = K
        Wm[n:, 3:] = p.T
        rhs = np.concatenate([t[b], np.zeros((3, 2))], 0)
        T = np.linalg.solve(Wm, rhs).T  # [2, 66]
        for r0 in range(0, Ho, 64):
            r1 = min(Ho, r0 + 64)
            X, Y = np.meshgrid(xt, yt[r0:r1])
            d2g = (X[None] - s[b, :, 0, None, None]) ** 2 + (Y[None] - s[b, :, 1, None, None]) ** 2
            rg = d2g * np.log(d2g + 1e-6)
            xs = T[0, 0] + T[0, 1] * X + T[0, 2] * Y + np.tensordot(T[0, 3:], rg, 1)
            ys = T[1, 0] + T[1, 1] * X + T[1, 2] * Y + np.tensordot(T[1, 3:], rg, 1)
            out_x[b, r0:r1] = (xs + 1.0) * W_in / 2.0
            out_y[b, r0:r1] = (ys + 1.0) * H_in / 2.0
    return out_x, out_y


# --------------------------------------------------------------------------------------
# correlation layers
# --------------------------------------------------------------------------------------
def cost_volume(x1, x2, sr):
    """cv[d,y,x] = leaky_relu_0.1(mean_c x1[c,y,x] * x2pad[c,y+j,x+i]), d = j*(2sr+1)+i.
    spatial_network.py:333-358 (norm=False), temporal_network.py:149-174."""
    bs, c, h, w = x1.shape
    xp = F.pad(x2, [sr] * 4)
    k = 2 * sr + 1
    out = torch.empty(bs, k * k, h, w)
    for j in range(k):
        for i in range(k):
            out[:, j * k + i] = (x1 * xp[:, :, j:j + h, i:i + w]).mean(1)
    return F.leaky_relu(out, 0.1)


def ccl(f1, f2):
    """Contextual correlation: spatial_network.py:369-425.  -> [bs,2,h,w] (flow_w, flow_h)."""
    bs, c, h, w = f1.shape
    n1 = F.normalize(f1, p=2, dim=1)
    n2 = F.normalize(f2, p=2, dim=1)
    n2p = F.pad(n2, [1, 1, 1, 1])
    # filter k = h2*w + w2 is the 3x3 patch of n2 centred at (h2, w2), layout [c,3,3]
    patches = n2p.unfold(2, 3, 1).unfold(3, 3, 1)  # [bs,c,h,w,3,3]
    filt = patches.permute(0, 2, 3, 1, 4, 5).reshape(bs, h * w, c, 3, 3)
    flows = []
    kk = torch.arange(h * w, dtype=torch.float32)
    kh = torch.floor(kk / w)[:, None, None]
    kw = (kk - torch.floor(kk / w) * w)[:, None, None]
    hh = torch.arange(h, dtype=torch.float32)[None, :, None]
    ww = torch.arange(w, dtype=torch.float32)[None, None, :]
    for b in range(bs):
        match = F.conv2d(n1[b:b + 1], filt[b], padding=1)[0]  # [hw, h, w]
        p = F.softmax(match * 10, 0)
        fh = (p * (kh - hh)).sum(0, keepdim=True)
        fw = (p * (kw - ww)).sum(0, keepdim=True)
        flows.append(torch.cat([fw, fh], 0))
    return torch.stack(flows, 0)


# --------------------------------------------------------------------------------------
# networks from a state_dict
# --------------------------------------------------------------------------------------
def _bn(x, sd, pfx):
    return F.batch_norm(x, sd[pfx + ".running_mean"], sd[pfx + ".running_var"],
                        sd[pfx + ".weight"], sd[pfx + ".bias"], False, 0.0, 1e-5)


def _basic_block(x, sd, pfx, stride):
    """torchvision BasicBlock, BN in eval mode."""
    idt = x
    o = F.conv2d(x, sd[pfx + ".conv1.weight"], None, stride, 1)
    o = F.relu(_bn(o, sd, pfx + ".bn1"))
    o = F.conv2d(o, sd[pfx + ".conv2.weight"], None, 1, 1)
    o = _bn(o, sd, pfx + ".bn2")
    if (pfx + ".downsample.0.weight") in sd:
        idt = _bn(F.conv2d(x, sd[pfx + ".downsample.0.weight"], None, stride, 0), sd, pfx + ".downsample.1")
    return F.relu(o + idt)


def resnet_stage1(x, sd, pfx="feature_extractor_stage1"):
    """conv1,bn1,relu,maxpool,layer1,layer2 -> [bs,128,H/8,W/8].  spatial_network.py:123-139."""
    o = F.conv2d(x, sd[pfx + ".0.weight"], None, 2, 3)
    o = F.relu(_bn(o, sd, pfx + ".1"))
    o = F.max_pool2d(o, 3, 2, 1)
    o = _basic_block(o, sd, pfx + ".4.0", 1)
    o = _basic_block(o, sd, pfx + ".4.1", 1)
    o = _basic_block(o, sd, pfx + ".5.0", 2)
    o = _basic_block(o, sd, pfx + ".5.1", 1)
    return o


def resnet_stage2(x, sd, pfx="feature_extractor_stage2"):
    """layer3 -> [bs,256,H/16,W/16]."""
    o = _basic_block(x, sd, pfx + ".0.0", 2)
    return _basic_block(o, sd, pfx + ".0.1", 1)


def _regress_convs(x, sd, pfx, conv_ids, pool_after):
    for i in conv_ids:
        x = F.relu(F.conv2d(x, sd["%s.%d.weight" % (pfx, i)], None, 1, 1))
        if i in pool_after:
            x = F.max_pool2d(x, 2, 2)
    return x


def _regress_fc(x, sd, pfx):
    x = F.relu(F.linear(x, sd[pfx + ".0.weight"], sd[pfx + ".0.bias"]))
    x = F.relu(F.linear(x, sd[pfx + ".2.weight"], sd[pfx + ".2.bias"]))
    return F.linear(x, sd[pfx + ".4.weight"], sd[pfx + ".4.bias"])


R1_CONVS, R1_POOLS = (0, 2, 5, 7, 10, 12), (2, 7, 12)            # spatial_network.py:147-168
R2_CONVS, R2_POOLS = (0, 2, 5, 7, 10, 12, 15, 17), (2, 7, 12, 17)  # :181-209


def spatial_forward(sd, img1, img2):
    """SpatialNet.forward, spatial_network.py:276-331 -> (offset_1[bs,8], off2_ref[bs,126],
    off2_tgt[bs,126])."""
    bs, _, H, W = img1.shape
    f1_64 = resnet_stage1(img1, sd)
    f1_32 = resnet_stage2(f1_64, sd)
    f2_64 = resnet_stage1(img2, sd)
    f2_32 = resnet_stage2(f2_64, sd)
    corr = ccl(f1_32, f2_32)
    t = _regress_convs(corr, sd, "regressNet1_part1", R1_CONVS, R1_POOLS)
    offset_1 = _regress_fc(t.reshape(bs, -1), sd, "regressNet1_part2")
    _, H_ref, H_tgt = homography_split(offset_1.reshape(-1, 4, 2), H, W, scale=8.0)
    w8, h8 = W / 8, H / 8
    M = torch.tensor([[w8 / 2.0, 0.0, w8 / 2.0], [0.0, h8 / 2.0, h8 / 2.0], [0.0, 0.0, 1.0]])
    Minv = torch.inverse(M)
    Hm_ref = torch.matmul(torch.matmul(Minv[None], H_ref), M[None])
    Hm_tgt = torch.matmul(torch.matmul(Minv[None], H_tgt), M[None])
    osz = (int(H / 8), int(W / 8))
    w1 = homo_warp(f1_64, Hm_ref, osz)
    w2 = homo_warp(f2_64, Hm_tgt, osz)
    t = _regress_convs(cost_volume(w1, w2, 5), sd, "regressNet2_part1_ref", R2_CONVS, R2_POOLS)
    off_ref = _regress_fc(t.reshape(bs, -1), sd, "regressNet2_part2_ref")
    t = _regress_convs(cost_volume(w2, w1, 5), sd, "regressNet2_part1_tgt", R2_CONVS, R2_POOLS)
    off_tgt = _regress_fc(t.reshape(bs, -1), sd, "regressNet2_part2_tgt")
    return offset_1, off_ref, off_tgt


def build_spatial(sd, img1, img2):
    """build_SpatialNet, spatial_network.py:63-118 -> motion1, motion2 [bs,7,9,2]."""
    bs, _, H, W = img1.shape
    o1, o_ref, o_tgt = spatial_forward(sd, img1, img2)
    return spatial_tail(o1, o_ref, o_tgt, H, W)


def spatial_tail(o1, o_ref, o_tgt, H, W):
    """The post-network part of build_SpatialNet (full-scale split + H2Mesh), :68-115."""
    bs = o1.shape[0]
    _, H_ref, H_tgt = homography_split(o1.reshape(-1, 4, 2), H, W, scale=1.0)
    rig = rigid_mesh(bs, H, W)
    mesh_ref = h2mesh(H_ref, rig) + o_ref.reshape(-1, GRID_H + 1, GRID_W + 1, 2)
    mesh_tgt = h2mesh(H_tgt, rig) + o_tgt.reshape(-1, GRID_H + 1, GRID_W + 1, 2)
    return mesh_ref - rig, mesh_tgt - rig


def temporal_forward(sd, frames):
    """build_TemporalNet, temporal_network.py:23-34,120-147: list of N [bs,3,H,W] ->
    list of N [bs,7,9,2]; element 0 is zeros; element k = motion of frame k w.r.t. k-1."""
    bs = frames[0].shape[0]
    out = [torch.zeros(bs, GRID_H + 1, GRID_W + 1, 2)]
    prev = resnet_stage1(frames[0], sd)
    for k in range(1, len(frames)):
        cur = resnet_stage1(frames[k], sd)
        t = _regress_convs(cost_volume(prev, cur, 3), sd, "regressNet2_part1", R2_CONVS, R2_POOLS)
        off = _regress_fc(t.reshape(bs, -1), sd, "regressNet2_part2")
        out.append(off.reshape(-1, GRID_H + 1, GRID_W + 1, 2))
        prev = cur
    return out


def smooth_forward(sd, tsmotion1, tsmotion2, smesh1, smesh2):
    """build_SmoothNet, smooth_network.py:23-41,64-157.  Inputs: 4 lists of T tensors
    [bs,7,9,2].  Returns the reference's 8-key dict, each [bs,T,7,9,2]."""
    def path(lst):
        acc = [lst[0]]
        for i in range(1, len(lst)):
            acc.append(acc[-1] + lst[i])
        return torch.stack(acc, 1)

    m1 = torch.stack(list(smesh1), 1)
    m2 = torch.stack(list(smesh2), 1)
    p1 = path(tsmotion1)
    p2 = path(tsmotion2)
    pf = "MotionPre."

    def emb(x, name):
        return F.relu(F.linear(x, sd[pf + name + ".0.weight"], sd[pf + name + ".0.bias"]))

    hid = torch.cat([emb(m1, "embedding1"), emb(p1, "embedding3"),
                     emb(m2, "embedding1"), emb(p2, "embedding3")], 4)  # bs,T,7,9,128
    x = hid.permute(0, 4, 1, 2, 3)
    for i in (0, 2, 4):
        x = F.relu(F.conv3d(x, sd[pf + "MotionConv3D.%d.weight" % i], sd[pf + "MotionConv3D.%d.bias" % i],
                            1, (2, 1, 1)))
    delta = F.linear(x.permute(0, 2, 3, 4, 1), sd[pf + "decoding.0.weight"], sd[pf + "decoding.0.bias"])
    d1, d2 = delta[..., 0:2], delta[..., 2:4]
    return dict(ori_path1=p1, smooth_path1=p1 + d1, ori_mesh1=m1, smooth_mesh1=m1 - d1,
                ori_path2=p2, smooth_path2=p2 + d2, ori_mesh2=m2, smooth_mesh2=m2 - d2)


# --------------------------------------------------------------------------------------
# sequence glue (test_online_tra.py)
# --------------------------------------------------------------------------------------
def tsmotion_prep(smotion, tmotion, img_h=360, img_w=480):
    """test_online_tra.py:309-347 for ONE view: lists of N [1,7,9,2] ->
    (smesh list, tsmotion list)."""
    rig = rigid_mesh(1, img_h, img_w)
    nrig = norm_mesh(rig, img_h, img_w)
    smesh, ts = [], []
    for k in range(len(smotion)):
        sm = rig + smotion[k]
        if k == 0:
            tsm = smotion[k] * 0
        else:
            sm_prev = rig + smotion[k - 1]
            tm = rig + tmotion[k]
            moved = tps_point(norm_mesh(tm, img_h, img_w), nrig, norm_mesh(sm_prev, img_h, img_w))
            tsm = recover_mesh(moved, img_h, img_w) - sm
        smesh.append(sm)
        ts.append(tsm)
    return smesh, ts


def smooth_stream(sd, smesh1, smesh2, ts1, ts2, win=7):
    """Sliding-window loop, test_online_tra.py:359-392 -> smooth_mesh1/2 [1,N,7,9,2]."""
    S1 = S2 = None
    for k in range(len(smesh1) - (win - 1)):
        a1 = list(ts1[k:k + win])
        a2 = list(ts2[k:k + win])
        a1[0] = a1[0] * 0
        a2[0] = a2[0] * 0
        o = smooth_forward(sd, a1, a2, smesh1[k:k + win], smesh2[k:k + win])
        if k == 0:
            S1, S2 = o["smooth_mesh1"], o["smooth_mesh2"]
        else:
            S1 = torch.cat([S1, o["smooth_mesh1"][:, -1:]], 1)
            S2 = torch.cat([S2, o["smooth_mesh2"][:, -1:]], 1)
    return S1, S2


def canvas(smooth_mesh1, smooth_mesh2, img_h, img_w):
    """test_online_tra.py:103-120: rescale to hr pixels, global min/max.
    Returns (mesh1_hr, mesh2_hr, width_min, height_min, out_width, out_height) with fp32
    0-dim tensors for the four scalars."""
    m1 = torch.stack([smooth_mesh1[..., 0] * img_w / 480, smooth_mesh1[..., 1] * img_h / 360], 4)
    m2 = torch.stack([smooth_mesh2[..., 0] * img_w / 480, smooth_mesh2[..., 1] * img_h / 360], 4)
    wmax = torch.maximum(m1[..., 0].max(), m2[..., 0].max())
    wmin = torch.minimum(m1[..., 0].min(), m2[..., 0].min())
    hmax = torch.maximum(m1[..., 1].max(), m2[..., 1].max())
    hmin = torch.minimum(m1[..., 1].min(), m2[..., 1].min())
    return m1, m2, wmin, hmin, wmax - wmin, hmax - hmin


def average_blend(a, b):
    """test_online_tra.py:142."""
    s = a + b + 1e-6
    return a * (a / s) + b * (b / s)


def stable_frame(img1, img2, mesh1_hr, mesh2_hr, wmin, hmin, out_w, out_h, mode="NORMAL"):
    """One iteration of the get_stable_sqe loop, AVERAGE fusion (test_online_tra.py:127-142).
    img [1,3,H,W] fp32 0..255; mesh_hr [1,7,9,2].  Returns (fused [3,Ho,Wo], warps [2,3,Ho,Wo])."""
    _, _, H, W = img1.shape
    nrig = norm_mesh(rigid_mesh(1, H, W), H, W)
    t1 = torch.stack([mesh1_hr[..., 0] - wmin, mesh1_hr[..., 1] - hmin], 3)
    t2 = torch.stack([mesh2_hr[..., 0] - wmin, mesh2_hr[..., 1] - hmin], 3)
    n1 = norm_mesh(t1, out_h, out_w)
    n2 = norm_mesh(t2, out_h, out_w)
    warp = tps_warp(torch.cat([img1, img2], 0), torch.cat([n1, n2], 0), torch.cat([nrig, nrig], 0),
                    (int(out_h.int()), int(out_w.int())), mode)
    return average_blend(warp[0], warp[1]), warp


def gaussian_blur_21(x, sigma=20.0, ksize=21):
    """torchvision.transforms.GaussianBlur(kernel_size=(21,21), sigma=20) as the reference's linear_blender builds it
    (test_online_tra.py:35): taps exp(-0.5 (t/sigma)^2) at t = linspace(-10, 10, 21), normalised; the 2-D kernel is the
    outer product; `reflect` padding of 10; one depthwise conv2d.  torchvision (third party, 0.14.1 pinned by
    environment.yml:358, not vendored) - restated from its published functional.gaussian_blur.  x [B,C,H,W]."""
    half = (ksize - 1) * 0.5
    t = torch.linspace(-half, half, ksize)
    k1 = torch.exp(-0.5 * (t / sigma) ** 2)
    k1 = k1 / k1.sum()
    k2 = k1[:, None] * k1[None, :]
    C = x.shape[1]
    xp = F.pad(x, [ksize // 2] * 4, mode="reflect")
    return F.conv2d(xp, k2[None, None].expand(C, 1, ksize, ksize).contiguous(), groups=C)


def linear_blender(ref, tgt, ref_m, tgt_m, mask=False, clean=False):
    """test_online_tra.py:34-58 `linear_blender`.  ref, tgt [1,3,Ho,Wo]; ref_m, tgt_m [1,1,Ho,Wo] warped masks.
    clean=False is the reference verbatim: `torch.nonzero` of the warped masks, whose out-of-image values are
    rounding residues of four clamped taps (about half of them non-zero), so the centroids - and with them the blending
    ramp - depend on those residues.  clean=True is the residue-free semantics of the CUDA path: masks thresholded at
    0.5 first (its lattice resampler returns exactly 0 outside and ~1 inside)."""
    if clean:
        ref_m, tgt_m = (ref_m > 0.5).float(), (tgt_m > 0.5).float()
    r1, c1 = torch.nonzero(ref_m[0, 0], as_tuple=True)
    r2, c2 = torch.nonzero(tgt_m[0, 0], as_tuple=True)
    center1 = (r1.float().mean(), c1.float().mean())
    center2 = (r2.float().mean(), c2.float().mean())
    vec = (center2[0] - center1[0], center2[1] - center1[1])
    ovl = (ref_m * tgt_m).round()[:, 0].unsqueeze(1)
    ref_m_ = ref_m[:, 0].unsqueeze(1) - ovl
    r, c = torch.nonzero(ovl[0, 0], as_tuple=True)
    ovl_mask = torch.zeros_like(ref_m_)
    proj_val = (r - center1[0]) * vec[0] + (c - center1[1]) * vec[1]
    ovl_mask[ovl.bool()] = (proj_val - proj_val.min()) / (proj_val.max() - proj_val.min() + 1e-3)
    mask1 = (gaussian_blur_21(ref_m_ + (1 - ovl_mask) * ref_m[:, 0].unsqueeze(1)) * ref_m + ref_m_).clamp(0, 1)
    if mask:
        return mask1
    mask2 = (1 - mask1) * tgt_m
    return ref * mask1 + tgt * mask2


def stable_frame_linear(img1, img2, mesh1_hr, mesh2_hr, wmin, hmin, out_w, out_h, mode="NORMAL", clean=False):
    """One iteration of the get_stable_sqe loop, LINEAR fusion (test_online_tra.py:143-150): a ones channel rides
    through the resampler as the mask.  Returns fused [3,Ho,Wo]."""
    _, _, H, W = img1.shape
    nrig = norm_mesh(rigid_mesh(1, H, W), H, W)
    t1 = torch.stack([mesh1_hr[..., 0] - wmin, mesh1_hr[..., 1] - hmin], 3)
    t2 = torch.stack([mesh2_hr[..., 0] - wmin, mesh2_hr[..., 1] - hmin], 3)
    n1 = norm_mesh(t1, out_h, out_w)
    n2 = norm_mesh(t2, out_h, out_w)
    one = torch.ones_like(img1[:, :1])
    warp = tps_warp(torch.cat([torch.cat([img1, one], 1), torch.cat([img2, one], 1)], 0), torch.cat([n1, n2], 0),
                    torch.cat([nrig, nrig], 0), (int(out_h.int()), int(out_w.int())), mode)
    return linear_blender(warp[0:1, 0:3], warp[1:2, 0:3], warp[0:1, 3:4], warp[1:2, 3:4], clean=clean)[0]


def three_view_meshes(w12m1, w12m2, w23m1, w23m2, img_h, img_w):
    """Middle-plane alignment of two stitched pairs that share their middle view
    (test_online_tra_threeview.py:345-455).  Inputs: smooth meshes [1,N,7,9,2] @480x360 of pair (1,2) and
    pair (2,3); w12m2 and w23m1 belong to the same physical view.
    Returns (mesh1, middle, mesh3 [1,N,7,9,2] in provisional-canvas pixels, width_min, height_min, out_width,
    out_height of the NEW canvas, 0-dim fp32 tensors)."""
    def hr(m):
        return torch.stack([m[..., 0] * img_w / 480, m[..., 1] * img_h / 360], 4)
    a1, a2, b1, b2 = hr(w12m1), hr(w12m2), hr(w23m1), hr(w23m2)
    # rigid alignment of pair (2,3) onto pair (1,2) by the mean vertex offset of the shared view (:354-360)
    off = (a2 - b1).reshape(a2.shape[0], a2.shape[1], -1, 2).mean(2)[:, :, None, None, :]
    b1, b2 = b1 + off, b2 + off
    mid = (a2 + b1) / 2.0                                                            # :363
    xs = [t[..., 0] for t in (a1, a2, b1, b2)]
    ys = [t[..., 1] for t in (a1, a2, b1, b2)]
    wmin, wmax = min(t.min() for t in xs), max(t.max() for t in xs)                  # :366-397
    hmin, hmax = min(t.min() for t in ys), max(t.max() for t in ys)
    ow, oh = wmax - wmin, hmax - hmin

    def sh(m):
        return torch.stack([m[..., 0] - wmin, m[..., 1] - hmin], 4)
    a1, a2, b1, b2, mid = sh(a1), sh(a2), sh(b1), sh(b2), sh(mid)                    # :406-410
    m1, m3 = [], []
    for i in range(mid.shape[1]):                                                    # :415-429
        n_a1, n_a2 = norm_mesh(a1[:, i], oh, ow), norm_mesh(a2[:, i], oh, ow)
        n_b1, n_b2 = norm_mesh(b1[:, i], oh, ow), norm_mesh(b2[:, i], oh, ow)
        n_mid = norm_mesh(mid[:, i], oh, ow)
        m1.append(recover_mesh(tps_point(n_a1, n_a2, n_mid), oh, ow))
        m3.append(recover_mesh(tps_point(n_b2, n_b1, n_mid), oh, ow))
    m1, m3 = torch.stack(m1, 1), torch.stack(m3, 1)
    xs = [t[..., 0] for t in (m1, mid, m3)]                                           # :436-455
    ys = [t[..., 1] for t in (m1, mid, m3)]
    wmin2, wmax2 = min(t.min() for t in xs), max(t.max() for t in xs)
    hmin2, hmax2 = min(t.min() for t in ys), max(t.max() for t in ys)
    return m1, mid, m3, wmin2, hmin2, wmax2 - wmin2, hmax2 - hmin2


def three_view_frame(img1, img2, img3, mesh1, mesh2, mesh3, wmin, hmin, out_w, out_h, mode="NORMAL"):
    """One iteration of the three-view warp + AVERAGE fusion loop (test_online_tra_threeview.py:470-490):
    fuse(1,2) then fuse(12,3).  img [1,3,H,W]; mesh [1,7,9,2] (one frame).  Returns [3,Ho,Wo]."""
    _, _, H, W = img1.shape
    nrig = norm_mesh(rigid_mesh(1, H, W), H, W)
    ns = [norm_mesh(torch.stack([m[..., 0] - wmin, m[..., 1] - hmin], 3), out_h, out_w) for m in (mesh1, mesh2, mesh3)]
    warp = tps_warp(torch.cat([img1, img2, img3], 0), torch.cat(ns, 0), torch.cat([nrig, nrig, nrig], 0),
                    (int(out_h.int()), int(out_w.int())), mode)
    return average_blend(average_blend(warp[0], warp[1]), warp[2])


def three_view_frame_linear(img1, img2, img3, mesh1, mesh2, mesh3, wmin, hmin, out_w, out_h, mode="NORMAL", clean=False):
    """test_online_tra_threeview.py:492-503: three-image warp with a ones channel, fuse(1,2) with linear_blender,
    mask12 = mask1 + mask2 - mask1*mask2, fuse(12,3).  Returns [3,Ho,Wo]."""
    _, _, H, W = img1.shape
    nrig = norm_mesh(rigid_mesh(1, H, W), H, W)
    ns = [norm_mesh(torch.stack([m[..., 0] - wmin, m[..., 1] - hmin], 3), out_h, out_w) for m in (mesh1, mesh2, mesh3)]
    one = torch.ones_like(img1[:, :1])
    warp = tps_warp(torch.cat([torch.cat([im, one], 1) for im in (img1, img2, img3)], 0), torch.cat(ns, 0),
                    torch.cat([nrig, nrig, nrig], 0), (int(out_h.int()), int(out_w.int())), mode)
    m1, m2, m3 = warp[0:1, 3:4], warp[1:2, 3:4], warp[2:3, 3:4]
    if clean:
        m1, m2, m3 = (m1 > 0.5).float(), (m2 > 0.5).float(), (m3 > 0.5).float()
    img12 = linear_blender(warp[0:1, 0:3], warp[1:2, 0:3], m1, m2, clean=clean)
    mask12 = m1 + m2 - m1 * m2
    return linear_blender(img12, warp[2:3, 0:3], mask12, m3, clean=clean)[0]


def nview_meshes(pairs, img_h, img_w):
    """N-view generalisation of the middle-plane alignment (BASELINE.json config 5; the reference defines N = 3 only,
    Full_model_inference/README.md:39 / test_online_tra_threeview.py:345-455, and this reduces to it for N = 3 -
    tests/test_oracle_golden.py::test_nview_reduces_to_three_view).  `pairs`: list of N-1 (meshA, meshB) smooth meshes
    [1,F,7,9,2] @480x360 of the stitched pairs (1,2), (2,3), ..; meshB of pair p and meshA of pair p+1 are two
    instances of the same physical view.
      1. every pair p >= 2 is shifted by the per-frame mean vertex offset that brings its meshA onto the (already
         shifted) meshB of pair p-1                                                     (:354-360, chained)
      2. shared view k sits on the middle plane of its two instances                       (:363)
      3. provisional canvas over all shifted pair meshes                                   (:366-410)
      4. the two OUTER views follow their pair's instance of the neighbouring shared view through the TPS that maps
         it onto that view's middle plane                                                  (:415-429)
      5. new canvas over the N final meshes                                                (:436-455)
    Returns (list of N meshes [1,F,7,9,2] in provisional-canvas pixels, wmin, hmin, out_w, out_h of the new canvas)."""
    def hr(m):
        return torch.stack([m[..., 0] * img_w / 480, m[..., 1] * img_h / 360], 4)
    P = [[hr(a), hr(b)] for a, b in pairs]
    for p in range(1, len(P)):
        off = (P[p - 1][1] - P[p][0]).reshape(P[p][0].shape[0], P[p][0].shape[1], -1, 2).mean(2)[:, :, None, None, :]
        P[p] = [P[p][0] + off, P[p][1] + off]
    mids = [(P[k - 1][1] + P[k][0]) / 2.0 for k in range(1, len(P))]
    flat = [m for pr in P for m in pr]
    wmin, wmax = min(t[..., 0].min() for t in flat), max(t[..., 0].max() for t in flat)
    hmin, hmax = min(t[..., 1].min() for t in flat), max(t[..., 1].max() for t in flat)
    ow, oh = wmax - wmin, hmax - hmin

    def sh(m):
        return torch.stack([m[..., 0] - wmin, m[..., 1] - hmin], 4)
    P = [[sh(a), sh(b)] for a, b in P]
    mids = [sh(m) for m in mids]
    first, last = [], []
    for i in range(mids[0].shape[1]):
        nm = lambda t: norm_mesh(t[:, i], oh, ow)  # noqa: E731
        first.append(recover_mesh(tps_point(nm(P[0][0]), nm(P[0][1]), nm(mids[0])), oh, ow))
        last.append(recover_mesh(tps_point(nm(P[-1][1]), nm(P[-1][0]), nm(mids[-1])), oh, ow))
    out = [torch.stack(first, 1)] + mids + [torch.stack(last, 1)]
    wmin2, wmax2 = min(t[..., 0].min() for t in out), max(t[..., 0].max() for t in out)
    hmin2, hmax2 = min(t[..., 1].min() for t in out), max(t[..., 1].max() for t in out)
    return out, wmin2, hmin2, wmax2 - wmin2, hmax2 - hmin2


def nview_frame(imgs, meshes, wmin, hmin, out_w, out_h, mode="NORMAL"):
    """N-image warp + sequential AVERAGE fusion fuse(..fuse(fuse(1,2),3)..,N) (test_online_tra_threeview.py:487-490 for
    N = 3).  imgs: N x [1,3,H,W]; meshes: N x [1,7,9,2] (one frame).  Returns [3,Ho,Wo]."""
    _, _, H, W = imgs[0].shape
    nrig = norm_mesh(rigid_mesh(1, H, W), H, W)
    ns = [norm_mesh(torch.stack([m[..., 0] - wmin, m[..., 1] - hmin], 3), out_h, out_w) for m in meshes]
    warp = tps_warp(torch.cat(list(imgs), 0), torch.cat(ns, 0), torch.cat([nrig] * len(imgs), 0),
                    (int(out_h.int()), int(out_w.int())), mode)
    f = warp[0]
    for k in range(1, len(imgs)):
        f = average_blend(f, warp[k])
    return f


def stitch_stream(sd_spatial, sd_temporal, sd_smooth, lr1, lr2, hr1, hr2, mode="NORMAL"):
    """The whole hot path for one video (test_online_tra.py:284-399, AVERAGE fusion).
    lr*: lists of [1,3,360,480] in [-1,1]; hr*: lists of [1,3,H,W] fp32 0..255.
    Returns dict(smotion1/2, tmotion1/2, smooth_mesh1/2, frames=list of [Ho,Wo,3] numpy)."""
    n = len(lr1)
    sm1, sm2 = [], []
    for k in range(n):
        a, b = build_spatial(sd_spatial, lr1[k], lr2[k])
        sm1.append(a)
        sm2.append(b)
    tm1 = temporal_forward(sd_temporal, lr1)
    tm2 = temporal_forward(sd_temporal, lr2)
    smesh1, ts1 = tsmotion_prep(sm1, tm1)
    smesh2, ts2 = tsmotion_prep(sm2, tm2)
    S1, S2 = smooth_stream(sd_smooth, smesh1, smesh2, ts1, ts2)
    H, W = hr1[0].shape[2:]
    m1, m2, wmin, hmin, ow, oh = canvas(S1, S2, H, W)
    frames = []
    for k in range(n):
        fused, _ = stable_frame(hr1[k], hr2[k], m1[:, k], m2[:, k], wmin, hmin, ow, oh, mode)
        frames.append(fused.numpy().transpose(1, 2, 0))
    return dict(smotion1=sm1, smotion2=sm2, tmotion1=tm1, tmotion2=tm2, tsmotion1=ts1, tsmotion2=ts2,
                smooth_mesh1=S1, smooth_mesh2=S2, frames=frames,
                canvas=(int(oh.int()), int(ow.int())))


# --------------------------------------------------------------------------------------
# deterministic synthetic workload - shared by tests and bench
# --------------------------------------------------------------------------------------
def synth_frame(t, v, H, W, coherent=True):
    """Band-limited noise frame, fp32 0..255, [1,3,H,W].  With coherent=True all frames of
    a view share the same texture shifted by a seeded integer random walk (so TemporalNet
    sees motion); the two views share the texture with a ~35% horizontal offset."""
    g = torch.Generator().manual_seed(1000)
    zh, zw = H // 16 + 8, (W // 16) * 2 + 8
    z = torch.randn(1, 3, zh, zw, generator=g)
    if not coherent:
        g2 = torch.Generator().manual_seed(1000 + 2 * t + v)
        z = torch.randn(1, 3, zh, zw, generator=g2)
    big = F.interpolate(z, size=(zh * 16, zw * 16), mode="bicubic", align_corners=False)
    gw = torch.Generator().manual_seed(77)
    walk = torch.randint(-2, 3, (4096, 2), generator=gw).cumsum(0)
    dy = 32 + int(walk[t % 4096, 0]) % 32
    dx = 32 + int(walk[t % 4096, 1]) % 32 + (int(0.35 * W) if v == 1 else 0)
    crop = big[:, :, dy:dy + H, dx:dx + W]
    return ((torch.tanh(crop) + 1.0) * 127.5).contiguous()


def lowres(hr, h=360, w=480):
    """hr 0..255 -> net input [-1,1] at 360x480 (bilinear)."""
    return F.interpolate(hr, size=(h, w), mode="bilinear", align_corners=False) / 127.5 - 1.0
