"""The tensor-core DQN / DDQN network (csrc/maze_net.cu, through the C ABI) against a plain PyTorch fp32 reference of
the same net (tests/net_reference.py, restated from /root/reference/agents/ddqn_agent.py:18-52,113-152; dropout off).
Tolerances: bf16 operands with fp32 accumulation -- forward and gradients within 2e-2 / 5e-2 of the tensor's largest
magnitude (measured: ~5e-3 / ~3e-2); the AdamW kernel itself, fed identical gradients, to 2e-6."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from net_reference import RefDQN, ddqn_loss, pack_windows  # noqa: E402


def _close(got, ref, tol, what):
    got, ref = got.float(), ref.float()
    scale = max(ref.abs().max().item(), 1e-6)
    err = (got - ref).abs().max().item()
    assert err <= tol * scale, f"{what}: max abs err {err:.4g} vs scale {scale:.4g} (tol {tol})"


def _close_norm(got, ref, tol, what):
    """Gradients: relative Frobenius error <= tol and no entry off by more than 3 tol of the largest magnitude.  bf16
    activations can flip a discrete choice of single samples (a LeakyReLU / ReLU mask at a pre-activation next to zero, a
    max-pool or arg-max tie), which moves a whole row of a weight gradient by that sample's share: a norm is the
    measure that sees this as the small perturbation it is."""
    got, ref = got.float(), ref.float()
    rel = ((got - ref).norm() / ref.norm().clamp_min(1e-12)).item()
    worst = ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()
    assert rel <= tol and worst <= 3 * tol, f"{what}: relative Frobenius error {rel:.4g}, worst entry {worst:.4g} of max (tol {tol})"


def _batch(n, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    win = (torch.rand(n, 3, 15, 15, generator=g) < 0.45).float().cuda()
    nwin = (torch.rand(n, 3, 15, 15, generator=g) < 0.45).float().cuda()
    vec, nvec = torch.rand(n, 6, generator=g).cuda(), torch.rand(n, 6, generator=g).cuda()
    action = torch.randint(0, 4, (n,), generator=g).to(torch.uint8).cuda()
    reward = (torch.rand(n, generator=g) - 0.5).cuda()
    return dict(vec=vec, win=win, pwin=pack_windows(win), nvec=nvec, nwin=nwin, pnwin=pack_windows(nwin), action=action, reward=reward)


def _refs(net):
    src, tgt = RefDQN().cuda(), RefDQN().cuda()
    src.load_state_dict(net.state_dict("source"))
    tgt.load_state_dict(net.state_dict("target"))
    return src, tgt


@pytest.fixture(autouse=True)
def _no_tf32():
    a, b = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = a, b


@pytest.mark.parametrize("M,N,K,tile,act", [(128, 256, 64, 256, 0), (130, 264, 200, 256, 1), (72, 8, 40, 128, 2), (1000, 1568, 1024, 256, 0),
                                            (384, 1024, 512, 128, 1),
                                            # tile 512: CTA pairs (cta_group::2), 256 x 256 tiles
                                            (256, 256, 64, 512, 0), (512, 512, 448, 512, 1), (1000, 1568, 1024, 512, 2), (8192, 1024, 1600, 512, 1),
                                            (130, 264, 200, 512, 0)])
def test_gemm_bias_activation(M, N, K, tile, act):
    from maze_b200.dqn_net import gemm_bf16
    torch.manual_seed(M + N + K)
    A, B = (torch.randn(M, K, device="cuda") * 0.5).bfloat16(), (torch.randn(N, K, device="cuda") * 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    C = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    gemm_bf16(A, B, C, 0, act, bias=bias, tile_n=tile)
    ref = A.float() @ B.float().t() + bias
    ref = torch.nn.functional.leaky_relu(ref, 0.01) if act == 1 else (torch.relu(ref) if act == 2 else ref)
    _close(C, ref, 1e-2, "gemm")


def test_gemm_mask_and_split_k_accumulate():
    from maze_b200.dqn_net import gemm_bf16
    torch.manual_seed(5)
    A, B = (torch.randn(200, 128, device="cuda") * 0.5).bfloat16(), (torch.randn(520, 128, device="cuda") * 0.5).bfloat16()
    aux = torch.randn(200, 520, device="cuda").bfloat16()
    C = torch.zeros(200, 520, device="cuda", dtype=torch.bfloat16)
    gemm_bf16(A, B, C, 1, 1, aux=aux, tile_n=128)
    _close(C, (A.float() @ B.float().t()) * torch.where(aux.float() > 0, 1.0, 0.01), 1e-2, "masked gemm")
    A, B = (torch.randn(512, 4096, device="cuda") * 0.5).bfloat16(), (torch.randn(1024, 4096, device="cuda") * 0.5).bfloat16()
    C = torch.ones(512, 1024, device="cuda")
    gemm_bf16(A, B, C, 2, 0, tile_n=256, splits=4)
    _close(C, A.float() @ B.float().t() + 1.0, 1e-4, "split-K accumulate")


@pytest.mark.parametrize("K,M,N,tile,splits", [(256, 128, 256, 256, 1), (1000, 512, 1024, 256, 4), (8192, 1024, 1600, 256, 4), (520, 72, 200, 128, 2)])
def test_gemm_transposed_operands(K, M, N, tile, splits):
    """C += A^T . B straight from [batch, features] activations (MN-major UMMA operands): the weight-gradient GEMMs."""
    from maze_b200.dqn_net import gemm_bf16
    torch.manual_seed(K + M)
    A, B = (torch.randn(K, M, device="cuda") * 0.5).bfloat16(), (torch.randn(K, N, device="cuda") * 0.5).bfloat16()
    C = torch.ones(M, N, device="cuda")
    gemm_bf16(A, B, C, 3, 0, tile_n=tile, splits=splits)
    _close(C, A.float().t() @ B.float() + 1.0, 1e-4, "A^T B accumulate")


def test_features_match_conv_pool():
    from maze_b200.dqn_net import DQNNet
    net = DQNNet("cuda", max_batch=512, seed=1)
    src, _ = _refs(net)
    b = _batch(300, 2)
    X, idx = net.features(b["vec"], b["pwin"], save_idx=True)
    with torch.no_grad():
        pre = torch.nn.functional.conv2d(b["win"], src.conv[0].weight.bfloat16().float(), src.conv[0].bias, padding=1)
        fw = torch.nn.functional.max_pool2d(torch.nn.functional.leaky_relu(pre, 0.01), 2, 2).flatten(1)
    _close(X[:, :1568], fw, 1e-2, "conv features")
    _close(X[:, 1568:1574], b["vec"], 1e-2, "state vector")
    assert (X[:, 1574:] == 0).all()
    pre_p = pre[:, :, :14, :14].reshape(300, 32, 7, 2, 7, 2).permute(0, 1, 2, 4, 3, 5).reshape(300, 32 * 49, 4)
    picked = pre_p.gather(2, (idx & 3).long().unsqueeze(-1)).squeeze(-1)
    _close(picked, pre_p.max(-1)[0], 2e-2, "pool choice points at a maximum")
    assert ((((idx >> 2) & 1).bool() == (picked > 0)) | (picked.abs() < 2e-2)).all()


@pytest.mark.parametrize("n", [1, 37, 1000])
def test_forward_matches_reference(n):
    from maze_b200.dqn_net import DQNNet
    net = DQNNet("cuda", max_batch=1024, seed=3)
    src, _ = _refs(net)
    b = _batch(n, n)
    q = net.forward(b["vec"], b["pwin"])
    with torch.no_grad():
        ref = src((b["vec"], b["win"]))
    _close(q, ref, 2e-2, "forward")
    agree = (q.argmax(1) == ref.argmax(1)).float().mean().item()
    assert agree > 0.97, agree


@pytest.mark.parametrize("n", [256, 1000])
def test_backward_matches_autograd(n):
    from maze_b200.dqn_net import DQNNet, _views
    net = DQNNet("cuda", max_batch=1024, seed=4)
    g = torch.Generator().manual_seed(9)
    net.load_state_dict({k: v + 0.01 * torch.randn(v.shape, generator=g).to(v.device) for k, v in net.state_dict("source").items()}, which="target")
    src, tgt = _refs(net)
    b = _batch(n, 10 + n)
    net.grads.zero_()
    qsa = torch.zeros(n, device="cuda")
    net.backward(b["vec"], b["pwin"], b["nvec"], b["pnwin"], b["action"], b["reward"], 0.9, qsa_out=qsa)
    loss, rq = ddqn_loss(src, tgt, (b["vec"], b["win"]), b["action"], b["reward"], (b["nvec"], b["nwin"]), 0.9)
    loss.backward()
    _close(qsa, rq.detach(), 2e-2, "q(s, a)")
    _close(net.loss, loss.detach().view(1), 3e-2, "loss")
    gv = _views(net.grads)
    for name, p in src.named_parameters():
        got = gv[name]
        if name == "fc.0.weight":
            assert (got[:, 1574:] == 0).all()
            got = got[:, :1574]
        _close_norm(got.reshape(p.grad.shape), p.grad, 4e-2, f"grad {name}")


def test_adamw_kernel_matches_torch_on_identical_gradients():
    """The optimiser in isolation: feed the autograd gradients of the reference into net.grads."""
    from maze_b200.dqn_net import DQNNet, _views
    net = DQNNet("cuda", max_batch=256, seed=5)
    src, tgt = _refs(net)
    opt = torch.optim.AdamW(src.parameters(), 1e-3)
    for it in range(3):
        b = _batch(256, 20 + it)
        loss, _ = ddqn_loss(src, tgt, (b["vec"], b["win"]), b["action"], b["reward"], (b["nvec"], b["nwin"]), 0.9)
        opt.zero_grad()
        (loss * 300).backward()            # large enough for the +-1 clamp to bite on some entries
        gv = _views(net.grads)
        net.grads.zero_()
        for name, p in src.named_parameters():
            if name == "fc.0.weight":
                gv[name][:, :1574] = p.grad
            else:
                gv[name].copy_(p.grad)
        for p in src.parameters():
            p.grad.data.clamp_(-1, 1)
        opt.step()
        net.adamw(1e-3)
        after = net.state_dict("source")
        for name, p in src.named_parameters():
            _close(after[name], p.detach(), 2e-6, f"adamw it{it} {name}")
        assert (net.grads == 0).all()
        assert torch.equal(net.w1_bf16, _views(net.params)["fc.0.weight"].bfloat16())
        assert torch.equal(net.w2t_bf16, _views(net.params)["fc.2.weight"].t().bfloat16())


def test_train_steps_track_the_reference():
    """optimize_model end to end on a fixed batch, device net and fp32 reference stepped side by side from the same
    weights: the loss trajectories stay together, the loss falls, and the parameter deltas agree wherever the first
    gradient is well above bf16 noise (Adam's first steps are lr * sign(g): entries with |g| near zero may flip)."""
    from maze_b200.dqn_net import DQNNet
    net = DQNNet("cuda", max_batch=512, seed=6)
    src, tgt = _refs(net)
    lr = 1e-4
    opt = torch.optim.AdamW(src.parameters(), lr)
    b = _batch(512, 33)
    before = net.state_dict("source")
    dev_losses, ref_losses = [], []
    for it in range(12):
        net.train_step(b["vec"], b["pwin"], b["nvec"], b["pnwin"], b["action"], b["reward"], gamma=0.9, lr=lr)
        dev_losses.append(net.loss.item())
        loss, _ = ddqn_loss(src, tgt, (b["vec"], b["win"]), b["action"], b["reward"], (b["nvec"], b["nwin"]), 0.9)
        ref_losses.append(loss.item())
        opt.zero_grad()
        loss.backward()
        if it == 0:
            g0 = {k: p.grad.clone() for k, p in src.named_parameters()}
        for p in src.parameters():
            p.grad.data.clamp_(-1, 1)
        opt.step()
    for it, (a, r) in enumerate(zip(dev_losses, ref_losses)):
        assert abs(a - r) <= 0.1 * r + 1e-4, (it, dev_losses, ref_losses)
    assert dev_losses[-1] < 0.9 * dev_losses[0], dev_losses
    after = net.state_dict("source")
    for name, p in src.named_parameters():
        big = g0[name].abs() > 0.2 * g0[name].abs().max()
        d_got, d_ref = (after[name] - before[name])[big], (p.detach() - before[name])[big]
        assert ((d_got - d_ref).abs() <= 0.25 * d_ref.abs().max()).float().mean().item() > 0.97, name


def test_update_target_and_state_dict_round_trip():
    from maze_b200.dqn_net import DQNNet
    net = DQNNet("cuda", max_batch=128, seed=7)
    b = _batch(128, 1)
    net.train_step(b["vec"], b["pwin"], b["nvec"], b["pnwin"], b["action"], b["reward"], gamma=0.9, lr=1e-2)
    q_s, q_t = net.forward(b["vec"], b["pwin"], which=0), net.forward(b["vec"], b["pwin"], which=1)
    assert not torch.equal(q_s, q_t)
    net.update_target()
    assert torch.equal(net.forward(b["vec"], b["pwin"], which=1), q_s)
    other = DQNNet("cuda", max_batch=128, seed=99)
    other.load_state_dict(net.state_dict("source"))
    assert torch.equal(other.forward(b["vec"], b["pwin"]), q_s)
    ref = RefDQN()
    ref.load_state_dict({k: v.cpu() for k, v in net.state_dict("source").items()})   # the reference module's parameter names


def test_sample_packed_draws_the_same_transitions_as_sample():
    import maze_b200 as mb
    from maze_b200.dqn import DeviceReplay, unpack_windows
    env = mb.MazeVectorEnv(256, shape=(21, 21), enrich=True, num_mazes=8, seed=3)
    mem = DeviceReplay(env, 4096, seed=5)
    env.reset()
    mem.observe()
    g = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(6):
        a = torch.randint(0, 4, (256,), device="cuda", generator=g).to(torch.uint8)
        env.step(a)
        mem.push(a)
    (vec, win), action, reward, (nvec, nwin) = mem.sample(500)
    mem._draw -= 1
    pv, pw, pnv, pnw, pa, pr = mem.sample_packed(500)
    assert torch.equal(pv, vec) and torch.equal(pnv, nvec) and torch.equal(pa.long(), action) and torch.equal(pr, reward)
    assert torch.equal(unpack_windows(pw), win) and torch.equal(unpack_windows(pnw), nwin)
