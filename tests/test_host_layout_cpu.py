"""Host logic of the one-launch formulations: the motion encoder's window formulation (stylegan_v_b200/time_encoder.py::trajectory_slabs + the slab semantics of
stylegan_v_b200/dense.py::conv1d_slabs), emulated with torch ops on CPU: evaluating the two valid conv1d layers only on the slabs must give the
trajectory codes the full conv1d formulation (the reference's, layers.py:356-373 / motion.py:100-115) gathers.  The CUDA kernels themselves are
checked against the same formulation in tests/test_dense_gpu.py."""
import pytest
import torch
import torch.nn.functional as F

from stylegan_v_b200.time_encoder import MotionMappingNetwork, trajectory_slabs


def slabs_ref(src, base, P, conv):
    """What dense.conv1d_slabs computes: row (g, p) = the contiguous [k, C] window of `src` starting at element base[g] + p * C, against the
    conv1d weight re-ordered to [O, k * C]; gains, bias and leaky ReLU as EqualizedConv1d applies them."""
    C = src.shape[-1]
    O, _, k = conv.weight.shape
    flat = src.reshape(-1)
    if base is None:
        base = torch.arange(src.shape[0]) * ((P + k - 1) * C)
    rows = torch.stack([flat[int(b) + p * C: int(b) + (p + k) * C] for b in base for p in range(P)])
    wr = conv.weight.permute(0, 2, 1).reshape(O, k * C) * conv.weight_gain
    y = F.leaky_relu(rows @ wr.t() + conv.bias * conv.bias_gain, 0.2)
    return y.reshape(len(base), P, O)


@pytest.mark.parametrize('frames', [1, 3, 16])
def test_window_formulation_equals_full_trajectory(frames):
    torch.manual_seed(frames)
    enc = MotionMappingNetwork(z_dim=32, v_dim=32, kernel_size=11, motion_z_distance=16, time_enc_dim=16, max_num_frames=1024).double()
    B = 3
    L = enc.traj_len()
    t = torch.randint(0, 1000, (B, frames)).double()
    t[0, 0] = 0.0
    t[-1, -1] = 1023.0                                                # first and last admissible positions
    z = torch.randn(B, L, 32, dtype=torch.float64)
    k, d = 11, enc.motion_z_distance
    left = (t / d).floor().long()
    # the reference formulation: full conv1d over the sequence, then gather
    trajs = enc.conv(z.permute(0, 2, 1)).permute(0, 2, 1)
    rows = torch.arange(B).unsqueeze(1).expand(B, frames)
    want_l, want_r = trajs[rows, left].reshape(B * frames, -1), trajs[rows, left + 1].reshape(B * frames, -1)
    windows, left2, base = trajectory_slabs(left, B, frames, L, k, 32)
    assert windows == (frames * (k + 1) <= L - k + 1)
    if windows:
        assert base.shape == (B * frames,) and torch.equal(left2, left)          # admissible positions are not clamped
        y1 = slabs_ref(z, base, k + 1, enc.conv[0])
        y2 = slabs_ref(y1, None, 2, enc.conv[1])
        got_l, got_r = y2[:, 0], y2[:, 1]
    else:
        assert base.shape == (B,)
        y1 = slabs_ref(z, base, L - k + 1, enc.conv[0])
        tr = slabs_ref(y1, None, L - 2 * k + 2, enc.conv[1])
        assert torch.allclose(tr, trajs, atol=1e-10)
        got_l, got_r = tr[rows, left].reshape(B * frames, -1), tr[rows, left + 1].reshape(B * frames, -1)
    assert torch.allclose(got_l, want_l, atol=1e-10) and torch.allclose(got_r, want_r, atol=1e-10)


def test_window_positions_are_clamped_into_the_sequence():
    left = torch.tensor([[-3, 0], [70, 500]])
    windows, clamped, base = trajectory_slabs(left, 2, 2, 86, 11, 8)
    assert windows and clamped.min() == 0 and clamped.max() == 86 - 22
    assert int(base.max()) + (2 * 11) * 8 <= 2 * 86 * 8               # the last window ends inside the [B, L, C] buffer


def test_stacked_affine_layout_equals_per_layer_affines():
    """SynthesisNetwork.affine_layout (the column groups of the one-launch stacked affine product, stylegan_v_b200/dense.py::stacked_affine):
    evaluating group g as ws[:, order[g]] @ wcat[col[g]:col[g+1]].T must give every layer the style its own affine computes from the w row the
    reference feeds it (networks.py:350-357: block b, layer l reads ws[:, w_idx(b) + l])."""
    from stylegan_v_b200.synthesis import SynthesisNetwork
    torch.manual_seed(0)
    net = SynthesisNetwork(w_dim=64, img_resolution=32, channel_base=2048, channel_max=64, motion_z_dim=32, motion_v_dim=32, time_enc_dim=16).double()
    groups, order, layers, col = net.affine_layout()
    assert order == list(range(net.num_ws)) and all(c % 8 == 0 for c in col) and len(col) == len(order) + 1
    ws = torch.randn(3, net.num_ws, 64, dtype=torch.float64)
    wcat = torch.cat([l.affine.weight for l in layers]) * layers[0].affine.weight_gain
    bcat = torch.cat([l.affine.bias for l in layers])
    stacked = torch.cat([ws[:, order[g]] @ wcat[col[g]:col[g + 1]].t() + bcat[col[g]:col[g + 1]] for g in range(len(order))], dim=1)
    pieces = dict(zip([id(l) for l in layers], stacked.split([l.affine.weight.shape[0] for l in layers], dim=1)))
    w_idx = 0
    for res in net.block_resolutions:
        block = getattr(net, f'b{res}')
        for j, layer in enumerate(block.layers()):
            assert torch.allclose(pieces[id(layer)], layer.affine(ws[:, w_idx + j]), atol=1e-12), (res, j)
        w_idx += block.num_conv
    # and the network's own evaluation (library branch on CPU) agrees
    lib = net._all_styles(ws)
    assert all(torch.allclose(lib[id(l)], pieces[id(l)], atol=1e-12) for l in layers)


@pytest.mark.parametrize('H,p', [(17, 0), (16, 1), (9, 0), (12, 1)])
def test_stride2_weight_gradient_phases(H, p):
    """native_conv.stride2_phases: the weight gradient of a stride-2 3x3 convolution as four stride-1 weight gradients over the pixel-parity
    views of the input (what the discriminator's down layers issue on the GPU).  The stride-1 contraction dw[t] = sum g[y, x] * view[y + dy,
    x + dx] (zero outside the view — the TMA's out-of-bounds fill) is emulated with torch ops and compared with autograd through F.conv2d."""
    from stylegan_v_b200.native_conv import stride2_phases
    torch.manual_seed(H + p)
    N, I, O, k = 2, 3, 4, 3
    x = torch.randn(N, I, H, H, dtype=torch.float64)
    w = torch.zeros(O, I, k, k, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x, w, stride=2, padding=p)
    g = torch.randn_like(y)
    want, = torch.autograd.grad(y, w, g)
    oh, ow = g.shape[2:]
    dwt = torch.zeros(k * k, O, I, dtype=torch.float64)
    seen = []
    for a, c, offs, slots in stride2_phases(k, p):
        view = x[:, :, a::2, c::2]
        vh, vw = view.shape[2:]
        for (dy, dx), slot in zip(offs, slots):
            ys, xs = torch.arange(oh) + dy, torch.arange(ow) + dx
            oky, okx = (ys >= 0) & (ys < vh), (xs >= 0) & (xs < vw)
            sub = view[:, :, ys[oky]][:, :, :, xs[okx]]
            tmp = torch.zeros(N, I, oh, ow, dtype=torch.float64)                 # positions outside the view stay zero (the TMA's fill)
            tmp[:, :, oky.nonzero().squeeze(1)[:, None], okx.nonzero().squeeze(1)[None, :]] = sub
            dwt[slot] = torch.einsum('noyx,niyx->oi', g, tmp)
            seen.append(slot)
    assert sorted(seen) == list(range(k * k))                            # every tap exactly once
    got = dwt.reshape(k, k, O, I).permute(2, 3, 0, 1)
    assert torch.allclose(got, want, atol=1e-10)
