"""bf16-operand emulation of one minibatch of the native PPO update (taco_b200/csrc/taco_ppo.cu).  TEST INFRASTRUCTURE.

The reference computes PPO_ActorCritic.evaluate (IsaacGymEnvs/algorithms/nets_asymmetry.py:356-377) and the losses of PPO.update
(ppo_asymmetry.py:190-212) in fp32; ``taco_b200.ppo.ppo_update`` (PyTorch autograd) is the fp32 twin pinned to the reference's own
update (tests/golden/ppo_update*.npz).  The native kernels multiply bf16 operands and accumulate in fp32; this module restates that
arithmetic in torch -- the same formulas with every GEMM operand rounded to bf16 at the same points (activations, weights, the
LSTM's hidden state, the back-propagated pre-activation gradients) and manual back-propagation -- so that the kernels can be held
to a tight tolerance (the comparison with the fp32 reference is dominated by ReLU masks that flip where a pre-activation is within
bf16 rounding of zero: a few percent of gradient norm, which says nothing about the kernels).
"""
import math

import torch


def r(x):
    """Round to bf16 (round-to-nearest-even), keep fp32 storage."""
    return x.to(torch.bfloat16).to(torch.float32)


def _linears(mlp):
    return [m for m in mlp.layers if isinstance(m, torch.nn.Linear)]


def mlp_forward(x0, linears):
    """Returns (activations [X_0 .. X_{L-1}] as bf16-valued fp32, pre-activation of the last layer)."""
    xs = [r(x0)]
    for lin in linears[:-1]:
        z = xs[-1] @ r(lin.weight).T + lin.bias
        xs.append(r(torch.relu(z)))
    return xs, xs[-1] @ r(linears[-1].weight).T + linears[-1].bias


def mlp_backward(xs, linears, dz_last):
    """dz_last: bf16-valued gradient of the last pre-activation.  Returns ({layer index: (dW, db)}, d loss / d input as fp32)."""
    grads, dz = {}, dz_last
    dx = None
    for l in range(len(linears) - 1, -1, -1):
        grads[l] = (dz.T @ xs[l], dz.sum(dim=0))
        dx = dz @ r(linears[l].weight)
        if l > 0:
            dz = r(dx * (xs[l] > 0).to(dx.dtype))
    return grads, dx


def minibatch_grads(agent, obs, states, act, old_logp, adv, ret, clip, pi_coef, vf_coef, ent_coef):
    """Gradients of one minibatch the way the native kernels compute them.  Returns ({parameter name: gradient}, stats)."""
    with torch.no_grad():
        B = obs.shape[0]
        a_lin, c_lin = _linears(agent.actor_mlp), _linears(agent.critic_mlp)
        lstm = agent.critic_encoder.layers
        H = lstm.hidden_size
        # ---- forward
        xs_a, pre = mlp_forward(obs.reshape(B, -1), a_lin)
        mean = torch.tanh(pre)
        w_ih, w_hh = r(lstm.weight_ih_l0), r(lstm.weight_hh_l0)
        b = lstm.bias_ih_l0 + lstm.bias_hh_l0
        b_hi = r(b)
        bias = b_hi + r(b - b_hi)
        h = torch.zeros(B, H, device=obs.device)
        c = torch.zeros(B, H, device=obs.device)
        saved = []
        for t in range(states.shape[1]):
            x = r(states[:, t, :])
            g = h @ w_hh.T + x @ w_ih.T + bias
            gi, gf, gg, go = torch.sigmoid(g[:, :H]), torch.sigmoid(g[:, H:2 * H]), torch.tanh(g[:, 2 * H:3 * H]), torch.sigmoid(g[:, 3 * H:])
            c_new = gf * c + gi * gg
            h_new = r(go * torch.tanh(c_new))
            saved.append((h, x, r(gi), r(gf), r(gg), r(go), c, c_new))
            h, c = h_new, c_new
        xs_c, value = mlp_forward(h, c_lin)
        value = value.reshape(B)
        # ---- losses (ppo_asymmetry.py:190-221)
        lsd = 2.0 * agent.log_std
        inv_sd = torch.exp(-lsd)
        z = (act - mean) * inv_sd
        k = mean.shape[1]
        logp = -0.5 * (z * z).sum(dim=1) - lsd.sum() - 0.5 * k * math.log(2.0 * math.pi)
        log_ratio = logp - old_logp
        ratio = torch.exp(log_ratio)
        s1, s2 = adv * ratio, adv * torch.clamp(ratio, 1.0 - clip, 1.0 + clip)
        dl_dlogp = torch.where(s1 <= s2, -adv * ratio, torch.zeros_like(adv))
        gl = pi_coef * dl_dlogp / B
        dz_a = r(gl.unsqueeze(1) * z * inv_sd * (1.0 - mean * mean))
        dv = value - ret
        dz_c = r((vf_coef * 2.0 * dv / B).unsqueeze(1))
        stats = dict(surrogate=float((-torch.minimum(s1, s2)).mean()), value_loss=float((dv * dv).mean()),
                     kl=float(((ratio - 1.0) - log_ratio).mean()), mean=mean, value=value)
        grads = {"log_std": (gl.unsqueeze(1) * 2.0 * (z * z - 1.0)).sum(dim=0) + ent_coef * -2.0}
        # ---- backward
        ga, _ = mlp_backward(xs_a, a_lin, dz_a)
        for l, (dw, db) in ga.items():
            grads[f"actor_mlp.layers.{2 * l}.weight"], grads[f"actor_mlp.layers.{2 * l}.bias"] = dw, db
        gc, dh = mlp_backward(xs_c, c_lin, dz_c)
        for l, (dw, db) in gc.items():
            grads[f"critic_mlp.layers.{2 * l}.weight"], grads[f"critic_mlp.layers.{2 * l}.bias"] = dw, db
        d_wih, d_whh, d_b = torch.zeros_like(w_ih), torch.zeros_like(w_hh), torch.zeros_like(b)
        dc = torch.zeros(B, H, device=obs.device)
        for t in range(states.shape[1] - 1, -1, -1):
            h_prev, x, gi, gf, gg, go, c_prev, c_cur = saved[t]
            tc = torch.tanh(c_cur)
            dc = dc + dh * go * (1.0 - tc * tc)
            dg = r(torch.cat((dc * gg * gi * (1.0 - gi), dc * c_prev * gf * (1.0 - gf), dc * gi * (1.0 - gg * gg), dh * tc * go * (1.0 - go)), dim=1))
            dc = dc * gf
            d_whh += dg.T @ h_prev
            d_wih += dg.T @ x
            d_b += dg.sum(dim=0)
            dh = dg @ w_hh
        grads["critic_encoder.layers.weight_ih_l0"], grads["critic_encoder.layers.weight_hh_l0"] = d_wih, d_whh
        grads["critic_encoder.layers.bias_ih_l0"], grads["critic_encoder.layers.bias_hh_l0"] = d_b, d_b.clone()
        return grads, stats
