"""GPU: the ASSEMBLED autoencoder training step against four consecutive calls of the reference's own
`AutoencoderTrainingWrapper.training_step` (golden written by oracle/gen_golden_training.py from the unmodified wrapper, CPU fp32):
generator / discriminator alternation, loss weights and names, warm-up switch, AuralossLoss argument order, AdamW per group.

Tolerances (stated): the engines run bf16 activations, the golden is fp32 — loss terms within 3 % (+1e-3 abs), discriminator hinge loss
within 1e-3 abs, the first AdamW update of each parameter group reproduces torch's formula on our own gradients (config lr / betas / weight decay).  Parameter
gradients are compared with the golden for information only - see the comment in the step test for why (the reference's loss gradient
itself turns by cos 0.54 under a 0.1 % perturbation of the reconstruction at this test point)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ae_training_step.npz")


def _cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))


def test_generator_loss_terms_on_a_fixed_reconstruction():
    """Every term of the generator loss, value and gradient w.r.t. the decoded audio, on the golden step-0 reconstruction (the
    wrapper's own loss modules produced the reference): isolates the loss kernels from the autoencoder backward."""
    from b200sat.discriminator import EncodecDiscriminatorTrain
    from b200sat.stft_loss import SumAndDifferenceSTFTLoss, autoencoder_mrstft_terms
    G = np.load(GOLD, allow_pickle=False)
    meta = json.loads(str(G["meta"]))
    dev = "cuda"
    reals = torch.from_numpy(G["reals"]).to(dev)
    dec0 = torch.from_numpy(G["diag.decoded"]).to(dev)
    stft = SumAndDifferenceSTFTLoss(sample_rate=44100, **meta["loss_config"]["spectral"]["config"])
    dsd = {k[len("disc_init."):]: torch.from_numpy(G[k]) for k in G.files if k.startswith("disc_init.")}
    disc = EncodecDiscriminatorTrain(dsd, n_ffts=tuple(meta["disc_fft"]), hop_lengths=tuple(meta["disc_hop"]), device=dev)

    def grad_of(fn):
        leaf = dec0.clone().requires_grad_(True)
        val = fn(leaf)
        (g,) = torch.autograd.grad(val, leaf)
        return float(val), g

    results = {}
    for i, name in enumerate(("mrstft", "left", "right")):
        results[name] = grad_of(lambda d, i=i: autoencoder_mrstft_terms(stft, d, reals)[i])
    results["adv"] = grad_of(lambda d: disc.generator_terms(reals, d)[0])
    results["fm"] = grad_of(lambda d: disc.generator_terms(reals, d)[1])
    bad = []
    for name, (val, g) in results.items():
        ref_v, ref_g = float(G[f"diag.value.{name}"]), torch.from_numpy(G[f"diag.grad.{name}"]).to(dev)
        c = _cos(g, ref_g)
        nr = float(g.norm() / ref_g.norm())
        print(f"\n[T2 term {name:7s}] value ours {val:+.6f} reference {ref_v:+.6f} | grad cos {c:.5f} norm ratio {nr:.4f}")
        # the hinge-free adversarial term's gradient passes through five bf16 LeakyReLU conv layers: 0.995; the others 0.999
        if abs(val - ref_v) > 2e-3 * abs(ref_v) + 1e-4 or c < (0.995 if name == "adv" else 0.999) or abs(nr - 1) > 2e-2:
            bad.append(name)
    assert not bad, bad
    # all five terms in ONE graph with the reference's weights: the accumulated gradient must equal the weighted sum of the parts
    leaf = dec0.clone().requires_grad_(True)
    sd_, l_, r_ = autoencoder_mrstft_terms(stft, leaf, reals)
    adv, fm = disc.generator_terms(reals, leaf)
    (1.0 * sd_ + 0.5 * l_ + 0.5 * r_ + 0.1 * adv + 5.0 * fm).backward()
    ref_total = sum(w * torch.from_numpy(G[f"diag.grad.{n}"]).to(dev) for n, w in (("mrstft", 1.0), ("left", 0.5), ("right", 0.5), ("adv", 0.1), ("fm", 5.0)))
    c = _cos(leaf.grad, ref_total)
    print(f"[T2 all terms in one graph] grad cos {c:.5f} norm ratio {float(leaf.grad.norm() / ref_total.norm()):.4f}")
    assert c >= 0.999


def test_four_steps_match_the_reference_wrapper():
    from oracle import oobleck as oo
    from b200sat.ae_training import AutoencoderTrainingStep
    from b200sat.autoencoder_train import OobleckTrainModel
    from b200sat.discriminator import EncodecDiscriminatorTrain
    G = np.load(GOLD, allow_pickle=False)
    meta = json.loads(str(G["meta"]))
    dev = "cuda"
    sd = oo.make_state_dict(channels=64, c_mults=(1, 2, 4), strides=(2, 4, 4), enc_latent=128, dec_latent=64, seed=meta["ae_weights_seed"])
    ae = OobleckTrainModel(sd, strides=(2, 4, 4), device=dev)
    dsd = {k[len("disc_init."):]: torch.from_numpy(G[k]) for k in G.files if k.startswith("disc_init.")}
    disc = EncodecDiscriminatorTrain(dsd, n_ffts=tuple(meta["disc_fft"]), hop_lengths=tuple(meta["disc_hop"]), device=dev)
    step = AutoencoderTrainingStep(ae, disc, loss_config=meta["loss_config"], optimizer_configs=meta["optimizer_configs"], sample_rate=44100,
                                   warmup_steps=0, use_ema=False)
    reals = torch.from_numpy(G["reals"]).to(dev)
    ae_named = {n: getattr(ae, n.replace(".", "__")) for n in ae.names}
    disc_named = {n: getattr(disc, n.replace(".", "__")) for n in disc.names}
    for s in range(4):
        is_d = s % 2 == 1
        named = disc_named if is_d else ae_named
        watch = sorted(k[len(f"s{s}.grad."):] for k in G.files if k.startswith(f"s{s}.grad."))
        before = {n: named[n].detach().clone() for n in watch}
        noise = torch.from_numpy(G[f"s{s}.vae_noise"]).to(dev)
        loss, log = step.training_step(reals, vae_noise=noise)
        ref_loss = float(G[f"s{s}.loss"])
        print(f"\n[T2 step {s} {'D' if is_d else 'G'}] loss ours {float(loss):.5f} reference {ref_loss:.5f}")
        # steps 0 / 1 start from identical weights; steps 2 / 3 follow one sign-like Adam update per group computed from bf16 gradients,
        # so the two trajectories have drifted apart: 3 % there becomes 10 %
        tol = 3e-2 if s < 2 else 1e-1
        if is_d:
            assert set(log) == {"train/disc_lr", "train/discriminator_loss"}
            assert abs(float(loss) - ref_loss) <= 1e-3 + (5e-3 if s < 2 else 2e-2) * abs(ref_loss)
        else:
            assert abs(float(loss) - ref_loss) <= tol * abs(ref_loss)
            for k in ("loss_adv", "feature_matching_loss", "mrstft_loss", "stft_loss_left", "stft_loss_right", "kl_loss"):
                ours, ref = float(log["train/" + k]), float(G[f"s{s}.log.{k}"])
                print(f"    {k:24s} ours {ours:+.5f} reference {ref:+.5f}")
                assert abs(ours - ref) <= tol * abs(ref) + 1e-3, (k, ours, ref)
            assert abs(float(log["train/gen_lr"]) - float(G[f"s{s}.log.gen_lr"])) < 1e-12
        # Parameter gradients vs the fp32 golden are PRINTED, not asserted: at this (random-init) test point the reference's own loss
        # gradient is ill-conditioned - perturbing the decoded audio by 0.1 % (fp32, reference auraloss code) already turns the MRSTFT
        # gradient by cos 0.54 (the log-magnitude term is dominated by near-empty bins), and our bf16 autoencoder output differs from the
        # fp32 one by 1.7 %.  What IS asserted is the decomposition: every loss term's gradient on a FIXED reconstruction (cos >= 0.999,
        # test above), the autoencoder backward on this configuration (tests/test_autoencoder_train_gpu.py, cos 0.999), and below the
        # optimizer wiring on our own gradients.
        cat_o, cat_r = [], []
        for n in watch:
            g_ref = torch.from_numpy(G[f"s{s}.grad.{n}"]).to(dev)
            g = named[n].grad
            cat_o.append(g.flatten().double()); cat_r.append(g_ref.flatten().double())
            print(f"    grad {n:60s} cos {_cos(g, g_ref):.4f} norm ratio {float(g.norm() / (g_ref.norm() + 1e-30)):.3f}")
        print(f"    all watched gradients concatenated: cos {_cos(torch.cat(cat_o), torch.cat(cat_r)):.4f}   (informational, see comment)")
        if s < 2:
            # first AdamW step of this group: m = (1-b1) g, v = (1-b2) g^2  =>  p' = p (1 - lr wd) - lr g / (|g| + eps)
            oc = meta["optimizer_configs"]["discriminator" if is_d else "autoencoder"]["optimizer"]["config"]
            lr, wd = oc["lr"], oc["weight_decay"]
            for n in watch:
                g = named[n].grad
                want = before[n] * (1 - lr * wd) - lr * g / (g.abs() + 1e-8)
                err = float((named[n].detach() - want).abs().max())
                assert err <= 2e-3 * lr + 1e-9, (n, err, lr)
    assert step.global_step == 4


def test_flat_parameters_keep_autograd_semantics():
    from b200sat.optim import FlatParameters, FusedAdamWEMA
    lin = torch.nn.Linear(8, 5).cuda()
    ref = torch.nn.Linear(8, 5).cuda()
    ref.load_state_dict(lin.state_dict())
    fp = FlatParameters(list(lin.parameters()))
    opt = FusedAdamWEMA(fp, lr=1e-2, betas=(0.8, 0.99), weight_decay=1e-3, ema=True, ema_before_step=True)
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-2, betas=(0.8, 0.99), weight_decay=1e-3)
    x = torch.randn(16, 8, device="cuda")
    w0 = lin.weight.detach().clone()
    for _ in range(3):
        opt.zero_grad(); ropt.zero_grad()
        lin(x).square().mean().backward(); ref(x).square().mean().backward()
        assert lin.weight.grad.data_ptr() == fp.flat_grad.data_ptr(), ".grad must stay a view of the flat buffer"
        opt.step(); ropt.step()
    assert torch.allclose(lin.weight, ref.weight, atol=1e-6) and torch.allclose(lin.bias, ref.bias, atol=1e-6)
    # EMA before step with update_after_step=1: first two updates copy the (pre-step) weights
    assert opt.ema is not None and torch.isfinite(opt.ema).all()
    assert not torch.allclose(lin.weight, w0)
