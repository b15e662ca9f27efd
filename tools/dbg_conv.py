import ctypes, sys, torch
sys.path.insert(0, ".")
from agilerl_b200 import _lib
from agilerl_b200.networks.spec import FlatLayout, rainbow_spec
lib = _lib.load()
def wgrad(first, obs_hw, c0, k0, s0, cout, k, s, rows):
    g = torch.Generator().manual_seed(1)
    spec = rainbow_spec((4, obs_hw, obs_hw), 3, channel_size=(c0, cout), kernel_size=(k0, k), stride_size=(s0, s),
                        latent_dim=16, hidden_size=(16,), obs_low=0.0, obs_high=255.0, obs_u8=True)
    layout = FlatLayout(spec); desc = layout.desc
    li = 0 if first else 1
    L = desc.enc[li]
    if first:
        ring = torch.randint(0, 256, (64, L.in_c, L.in_h, L.in_w), dtype=torch.uint8, generator=g)
        idx = torch.randint(0, 64, (rows,), generator=g)
        x64 = ring[idx].double() / 255.0
        x_dev, idx_dev = ring.cuda(), idx.cuda()
    else:
        x = torch.randn(rows, L.in_c, L.in_h, L.in_w, generator=g).relu()
        x64 = x.double(); x_dev, idx_dev = x.cuda(), None
    gout = torch.randn(rows, L.out_c, L.out_h, L.out_w, generator=g) * 0.1
    w64 = torch.zeros(L.out_c, L.in_c, L.ksize, L.ksize, dtype=torch.float64, requires_grad=True)
    b64 = torch.zeros(L.out_c, dtype=torch.float64, requires_grad=True)
    torch.nn.functional.conv2d(x64, w64, b64, stride=L.stride).backward(gout.double())
    grads = torch.full((layout.n_params,), float("nan"), dtype=torch.float32, device="cuda")
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    gdev = gout.cuda()
    n_st = lib.b2rl_conv_path_count(2)
    _lib.check(lib.b2rl_encoder_layer_wgrad(ctypes.byref(desc), li, x_dev.data_ptr(), idx_dev.data_ptr() if idx_dev is not None else None,
                                            rows, gdev.data_ptr(), grads.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(torch.device("cuda:0"))))
    torch.cuda.synchronize()
    gh = grads.cpu().double()
    gw = gh[L.w_off:L.w_off + w64.numel()].reshape(w64.shape); gb = gh[L.b_off:L.b_off + L.out_c]
    ref = w64.grad
    print("wgrad", first, obs_hw, cout, k, rows, "staged", lib.b2rl_conv_path_count(2) - n_st,
          "| dW err", (gw - ref).abs().max().item(), "scale", ref.abs().max().item(), "got max", gw.abs().max().item(),
          "| db err", (gb - b64.grad).abs().max().item(), "scale", b64.grad.abs().max().item())
    if (gw - ref).abs().max().item() > 1e-4 * ref.abs().max().item():
        ratio = (gw.reshape(L.out_c, -1) / ref.reshape(L.out_c, -1))
        print("   ratio[co=1, taps 0..15]", ratio[1, :16].tolist())
        print("   got[co=1,:8]", gw.reshape(L.out_c, -1)[1, :8].tolist(), "ref", ref.reshape(L.out_c, -1)[1, :8].tolist())
for a in [(False, 84, 32, 8, 4, 32, 4, 2, 256), (False, 84, 32, 8, 4, 32, 4, 2, 1), (True, 84, 32, 8, 4, 32, 4, 2, 5), (False, 26, 2, 4, 2, 8, 4, 2, 300)]:
    wgrad(*a)


def fwd(obs_hw, c0, k0, s0, cout, k, s, rows):
    g = torch.Generator().manual_seed(2)
    spec = rainbow_spec((4, obs_hw, obs_hw), 3, channel_size=(c0, cout), kernel_size=(k0, k), stride_size=(s0, s),
                        latent_dim=16, hidden_size=(16,), obs_low=0.0, obs_high=255.0, obs_u8=True)
    layout = FlatLayout(spec); desc = layout.desc
    L = desc.enc[1]
    hw = L.in_h
    params = torch.zeros(layout.n_params)
    w = torch.randn(cout, c0, k, k, generator=g) * (1.0 / (c0 * k * k) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    params[L.w_off:L.w_off + w.numel()] = w.reshape(-1)
    params[L.b_off:L.b_off + cout] = b
    x = torch.randn(rows, c0, hw, hw, generator=g).relu()
    ref = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), stride=s).relu()
    out = torch.full(ref.shape, float("nan"), dtype=torch.float32, device="cuda")
    ws = torch.empty(16 << 20, dtype=torch.uint8, device="cuda")
    pd, xd = params.cuda(), x.cuda()
    n_st = lib.b2rl_conv_path_count(2)
    _lib.check(lib.b2rl_encoder_layer_forward(ctypes.byref(desc), 1, pd.data_ptr(), xd.data_ptr(), None, rows,
                                              out.data_ptr(), ws.data_ptr(), ws.numel(), 0, _lib.stream_ptr(torch.device("cuda:0"))))
    torch.cuda.synchronize()
    got = out.cpu().double()
    err = (got - ref).abs()
    print("fwd", obs_hw, c0, cout, k, s, rows, "staged", lib.b2rl_conv_path_count(2) - n_st, "err", err.max().item(), "scale", ref.abs().max().item(),
          "nan", torch.isnan(got).sum().item(), "bad", (err > 1e-4).sum().item(), "of", err.numel())
    if err.max().item() > 1e-4:
        idx = (err > 1e-4).nonzero()
        print("   first bad idx", idx[:6].tolist(), "last", idx[-3:].tolist())
        i0 = tuple(idx[0].tolist())
        print("   got", got[i0].item(), "ref", ref[i0].item())


for a in [(34, 8, 4, 2, 16, 8, 2, 9), (34, 8, 4, 2, 16, 8, 2, 64), (34, 8, 4, 2, 32, 8, 2, 9), (34, 16, 4, 2, 16, 8, 2, 9)]:
    fwd(*a)
