"""The assembled autoencoder training step — `AutoencoderTrainingWrapper.training_step` (training/autoencoders.py:367-527) on the
b200sat engines: Oobleck VAE forward/backward (b200sat.autoencoder_train), EncodecDiscriminator losses and gradients
(b200sat.discriminator), the four MRSTFT terms in one pass (b200sat.stft_loss), fused AdamW (+ EMA) per parameter group.

What is kept from the reference, line by line:
  * `warmed_up` once `global_step >= warmup_steps` (:379-380); before that the adversarial / feature-matching terms are zero and, in
    `warmup_mode == "adv"`, the discriminator still trains (:436-447);
  * the step kind: discriminator step iff a discriminator exists, `global_step` is odd, and (`"full"` and warmed up, or `"adv"`)
    (:467-475); otherwise a generator step;
  * generator loss = sum of weight * term in this order (:162-243): adversarial `-mean D(fake)` (x 0.1), feature matching (x 5.0),
    sum/difference MRSTFT (x 1.0, with the AuralossLoss argument swap: input = reals, target = decoded,
    training/losses/losses.py:107-113), left and right MRSTFT (x 0.5 each), KL (x 1e-4); the logged values are the WEIGHTED terms
    under the reference's names;
  * the EMA of the autoencoder is updated BEFORE the generator's optimizer step (:499-500), with ema_pytorch's warm-up schedule;
  * `global_step` advances by one per training step (manual optimisation: one optimizer.step per call).
What is deliberately leaner (same parameter updates): the discriminator step runs the autoencoder without a tape (the reference
back-propagates `loss_dis` into the autoencoder and then discards those gradients with `opt_gen.zero_grad()` on the next generator
step), and the generator step does not build the discriminator-parameter graph."""
import torch

from .optim import FlatParameters, FusedAdamWEMA
from .stft_loss import SumAndDifferenceSTFTLoss, autoencoder_mrstft_terms


def _adamw(group, cfg, ema=False, ema_before_step=False):
    oc = cfg["optimizer"]
    if oc.get("type", "AdamW") != "AdamW":
        raise NotImplementedError(f"optimizer type {oc.get('type')!r}: only AdamW is implemented")
    c = oc["config"]
    return FusedAdamWEMA(group, lr=c["lr"], betas=tuple(c.get("betas", (0.9, 0.999))), eps=c.get("eps", 1e-8),
                         weight_decay=c.get("weight_decay", 1e-2), ema=ema, ema_before_step=ema_before_step)


class AutoencoderTrainingStep:
    def __init__(self, autoencoder, discriminator=None, loss_config=None, optimizer_configs=None, sample_rate=44100, warmup_steps=0,
                 warmup_mode="adv", use_ema=True, world_size=1, all_reduce=None):
        """autoencoder: b200sat.autoencoder_train.OobleckTrainModel; discriminator: b200sat.discriminator.EncodecDiscriminatorTrain or None;
        loss_config / optimizer_configs: the reference's `training.loss_configs` / `training.optimizer_configs` dictionaries
        (configs/model_configs/autoencoders/stable_audio_2_0_vae.json:41-116).  all_reduce(flat_grad_tensor): data-parallel hook."""
        self.ae, self.disc = autoencoder, discriminator
        lc = loss_config
        if lc["spectral"]["type"] != "mrstft":
            raise NotImplementedError("spectral loss type " + lc["spectral"]["type"])
        if lc.get("time", {}).get("weights", {}).get("l1", 0.0) or lc.get("time", {}).get("weights", {}).get("l2", 0.0):
            raise NotImplementedError("time-domain L1 / L2 terms")
        self.w_stft = float(lc["spectral"]["weights"]["mrstft"])
        self.w_kl = float(lc.get("bottleneck", {}).get("weights", {}).get("kl", 1e-6))
        self.use_disc = discriminator is not None
        if self.use_disc:
            self.w_adv = float(lc["discriminator"]["weights"]["adversarial"])
            self.w_fm = float(lc["discriminator"]["weights"]["feature_matching"])
        self.stft = SumAndDifferenceSTFTLoss(sample_rate=sample_rate, **lc["spectral"]["config"])
        self.warmup_steps, self.warmup_mode = int(warmup_steps), warmup_mode
        self.warmed_up = False
        self.global_step = 0
        self.world_size, self.all_reduce = world_size, all_reduce
        self.gen_params = FlatParameters(list(autoencoder.parameters()))
        self.opt_gen = _adamw(self.gen_params, optimizer_configs["autoencoder"], ema=use_ema, ema_before_step=True)
        if self.use_disc:
            self.disc_params = FlatParameters(list(discriminator.parameters()))
            self.opt_disc = _adamw(self.disc_params, optimizer_configs["discriminator"])

    def _reduce(self, group):
        if self.all_reduce is not None and self.world_size > 1:
            self.all_reduce(group.flat_grad)

    def training_step(self, reals, vae_noise=None):
        """reals fp32 [B, 2, T] on the device; vae_noise [B, latent, T/ratio] (None: drawn here, as `torch.randn_like` in
        models/bottleneck.py:105-134).  Returns (loss tensor, log dict with the reference's `train/...` keys)."""
        if self.global_step >= self.warmup_steps:
            self.warmed_up = True
        B = reals.shape[0]
        disc_step = bool(self.use_disc and self.global_step % 2
                         and ((self.warmup_mode == "full" and self.warmed_up) or self.warmup_mode == "adv"))
        log = {}
        scale = 1.0 / self.world_size
        if disc_step:
            with torch.no_grad():
                decoded = self.ae(reals, vae_noise)[0]
            loss = self.disc.discriminator_loss(reals, decoded)
            log["train/disc_lr"] = self.opt_disc.param_groups[0]["lr"]
            self.opt_disc.zero_grad()
            (loss * scale).backward()
            self._reduce(self.disc_params)
            self.opt_disc.step()
            log["train/discriminator_loss"] = loss.detach()
        else:
            decoded, kl, latents = self.ae(reals, vae_noise)
            terms = {}
            if self.use_disc:
                if self.warmed_up:
                    adv, fm = self.disc.generator_terms(reals, decoded)
                else:
                    adv = fm = torch.zeros((), device=reals.device)
                terms["loss_adv"] = self.w_adv * adv
                terms["feature_matching_loss"] = self.w_fm * fm
            sd, left, right = autoencoder_mrstft_terms(self.stft, decoded, reals)
            terms["mrstft_loss"] = self.w_stft * sd
            terms["stft_loss_left"] = (self.w_stft / 2) * left
            terms["stft_loss_right"] = (self.w_stft / 2) * right
            terms["kl_loss"] = self.w_kl * kl
            loss = sum(terms.values())
            self.opt_gen.zero_grad()
            (loss * scale).backward()
            self._reduce(self.gen_params)
            self.opt_gen.step()          # EMA (of the pre-step weights) + AdamW in one pass
            log["train/loss"] = loss.detach()
            log["train/latent_std"] = latents.std().detach()
            log["train/gen_lr"] = self.opt_gen.param_groups[0]["lr"]
            for k, v in terms.items():
                log["train/" + k] = v.detach()
        self.global_step += 1
        return loss.detach(), log
