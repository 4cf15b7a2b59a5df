"""Training objective of GeoLDM on padded batches (BASELINE config 5).

Mirrors the loss side of equivariant_diffusion/en_diffusion.py — ``EnVariationalDiffusion.compute_loss`` (:568-688),
``kl_prior`` (:406-435), ``compute_error`` (:451-459), ``log_constants_p_x_given_z0`` (:461-476),
``EnLatentDiffusion.forward`` (:1136-1191), ``log_constants_p_h_given_z0`` (:1083-1097),
``EnHierarchicalVAE.compute_loss`` / ``compute_reconstruction_error`` (:851-926) — and qm9/losses.py:12-43.
The heavy parts are the three EGNN evaluations (encoder without grad on the fused inference kernels; decoder and
denoiser with autograd through train.py); everything in this file is O(batch x atoms x 10) element-wise algebra.

Random draws: the reference pulls them from torch's global generator (randint for t, randn for the noises).  Every
function that draws takes an optional ``draws`` dict so that parity tests can inject the reference's own draws
(``eps_enc``, ``t_int``, ``eps_t``, ``eps_0``); without it draws come from ``torch.randn``/``torch.randint`` on the device.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

LOG_2PI = math.log(2.0 * math.pi)


def sum_except_batch(v):
    return v.reshape(v.shape[0], -1).sum(-1)


def remove_mean_with_mask(x, node_mask):
    n = node_mask.sum(1, keepdim=True)
    return x - (x.sum(1, keepdim=True) / n) * node_mask


def masked_noise(bs, n_nodes, n_dims, nf, node_mask):
    """CoM-free N(0,I) for the coordinates, plain N(0,I) for the features, both masked
    (equivariant_diffusion/utils.py:107-117,137-140; en_diffusion.py:702-714)."""
    zx = torch.randn(bs, n_nodes, n_dims, device=node_mask.device) * node_mask
    zx = remove_mean_with_mask(zx, node_mask)
    zh = torch.randn(bs, n_nodes, nf, device=node_mask.device) * node_mask
    return torch.cat([zx, zh], dim=2)


def gaussian_kl(q_mu, q_sigma, p_sigma_is_one_mask):
    """KL(N(q_mu, q_sigma) || N(0, 1)) summed over masked entries (en_diffusion.py:84-100 with p = N(0,1))."""
    term = torch.log(1.0 / (q_sigma + 1e-8) + 1e-8) + 0.5 * (q_sigma ** 2 + q_mu ** 2) - 0.5
    return sum_except_batch(term * p_sigma_is_one_mask)


def gaussian_kl_subspace(q_mu, q_sigma, d):
    """Same against N(0, 1) on a d-dimensional subspace with isotropic q_sigma [bs] (en_diffusion.py:103-120)."""
    mu2 = sum_except_batch(q_mu ** 2)
    return d * torch.log(1.0 / (q_sigma + 1e-8) + 1e-8) + 0.5 * (d * q_sigma ** 2 + mu2) - 0.5 * d


def _inflate(v, ndim):
    return v.reshape((v.shape[0],) + (1,) * (ndim - 1))


def _sigma(gamma, ndim=3):
    return _inflate(torch.sqrt(torch.sigmoid(gamma)), ndim)


def _alpha(gamma, ndim=3):
    return _inflate(torch.sqrt(torch.sigmoid(-gamma)), ndim)


def subspace_dimensionality(node_mask, n_dims):
    return (node_mask.squeeze(2).sum(1) - 1) * n_dims


def kl_prior(model, xh, node_mask):
    bs = xh.shape[0]
    gamma_T = model.gamma(torch.ones(bs, 1, device=xh.device))
    mu_T = _alpha(gamma_T) * xh
    mu_x, mu_h = mu_T[:, :, :model.n_dims], mu_T[:, :, model.n_dims:]
    sigma_h = _sigma(gamma_T)
    sigma_x = sigma_h.reshape(bs)
    kl_h = gaussian_kl(mu_h, sigma_h, node_mask)
    kl_x = gaussian_kl_subspace(mu_x, sigma_x, subspace_dimensionality(node_mask, model.n_dims))
    return kl_x + kl_h


def compute_error(model, net_out, eps):
    err = sum_except_batch((eps - net_out) ** 2)
    if model.training and model.loss_type == 'l2':
        err = err / ((model.n_dims + model.in_node_nf) * net_out.shape[1])
    return err


def log_constants_x(model, node_mask):
    """log of the Gaussian normaliser of p(x | z0) on the (n-1)*3 subspace (:461-476)."""
    bs = node_mask.shape[0]
    dof = subspace_dimensionality(node_mask, model.n_dims)
    gamma_0 = model.gamma(torch.zeros(bs, 1, device=node_mask.device)).reshape(bs)
    return dof * (-0.5 * gamma_0 - 0.5 * LOG_2PI)


def log_constants_h(model, node_mask):
    """Latent subclass override (:1083-1097): n_nodes * n_dims degrees of freedom (the reference's own count)."""
    bs = node_mask.shape[0]
    dof = node_mask.squeeze(2).sum(1) * model.n_dims
    gamma_0 = model.gamma(torch.zeros(bs, 1, device=node_mask.device)).reshape(bs)
    return dof * (-0.5 * gamma_0 - 0.5 * LOG_2PI)


def diffusion_loss(model, z_x, z_h, node_mask, edge_mask, context, t0_always, draws=None):
    """EnVariationalDiffusion.compute_loss on the latent (x part z_x, feature part z_h)."""
    bs, n_nodes = z_x.shape[0], z_x.shape[1]
    dev = z_x.device
    T = model.T
    l2_train = model.training and model.loss_type == 'l2'
    lowest_t = 1 if t0_always else 0
    if draws is not None and 't_int' in draws:
        t_int = draws['t_int'].to(dev).float().reshape(bs, 1)
    else:
        t_int = torch.randint(lowest_t, T + 1, size=(bs, 1), device=dev).float()
    t_is_zero = (t_int == 0).float().reshape(bs)
    s, t = (t_int - 1) / T, t_int / T
    gamma_s, gamma_t = model.gamma(s), model.gamma(t)
    alpha_t, sigma_t = _alpha(gamma_t), _sigma(gamma_t)
    if draws is not None and 'eps_t' in draws:
        eps = draws['eps_t'].to(dev)
    else:
        eps = masked_noise(bs, n_nodes, model.n_dims, model.in_node_nf, node_mask)
    xh = torch.cat([z_x, z_h], dim=2)
    z_t = alpha_t * xh + sigma_t * eps
    net_out = model.phi(z_t, t, node_mask, edge_mask, context)
    error = compute_error(model, net_out, eps)
    if l2_train:
        snr_weight = torch.ones_like(error)
    else:
        snr_weight = (torch.exp(-(gamma_s - gamma_t)) - 1).reshape(bs)
    loss_t_pos = 0.5 * snr_weight * error
    neg_log_const = torch.zeros_like(error) if l2_train else -log_constants_x(model, node_mask)
    kl = kl_prior(model, xh, node_mask)
    if t0_always:
        gamma_0 = model.gamma(torch.zeros_like(s))
        if draws is not None and 'eps_0' in draws:
            eps_0 = draws['eps_0'].to(dev)
        else:
            eps_0 = masked_noise(bs, n_nodes, model.n_dims, model.in_node_nf, node_mask)
        z_0 = _alpha(gamma_0) * xh + _sigma(gamma_0) * eps_0
        net_0 = model.phi(z_0, torch.zeros_like(s), node_mask, edge_mask, context)
        loss_0 = 0.5 * compute_error(model, net_0, eps_0)
        loss = kl + T * loss_t_pos + neg_log_const + loss_0
    else:
        loss_0 = 0.5 * error                                   # -log p(z0-ish | z_t) up to constants (:1122-1133)
        loss_t = loss_0 * t_is_zero + (1 - t_is_zero) * loss_t_pos
        loss = kl + (loss_t if l2_train else (T + 1) * loss_t) + neg_log_const
    return loss, {'t': t_int.reshape(bs), 'loss_t': loss, 'error': error}


def reconstruction_error(vae, xh_rec, xh):
    """Squared error on positions and charges + cross entropy on atom types (:851-884)."""
    bs, n_nodes, _ = xh.shape
    nd, nc = vae.n_dims, vae.num_classes
    err = sum_except_batch((xh_rec[:, :, :nd] - xh[:, :, :nd]) ** 2)
    logits = xh_rec[:, :, nd:nd + nc].reshape(bs * n_nodes, nc)
    target = xh[:, :, nd:nd + nc].reshape(bs * n_nodes, nc).argmax(dim=1)
    err = err + F.cross_entropy(logits, target, reduction='none').reshape(bs, n_nodes).sum(1)
    if vae.include_charges:
        err = err + sum_except_batch((xh_rec[:, :, -1:] - xh[:, :, -1:]) ** 2)
    if vae.training:
        err = err / ((nd + vae.in_node_nf) * n_nodes)
    return err


def vae_loss(vae, x, h, node_mask, edge_mask, context, draws=None):
    """EnHierarchicalVAE.compute_loss (:892-926): first-stage training objective."""
    xh = torch.cat([x, h['categorical'], h['integer']], dim=2)
    bs, n_nodes = x.shape[0], x.shape[1]
    z_x_mu, z_x_sigma, z_h_mu, z_h_sigma = vae.encode(x, h, node_mask, edge_mask, context)
    ones_h = torch.ones_like(z_h_sigma)
    kl_h = gaussian_kl(z_h_mu, ones_h, node_mask)
    kl_x = gaussian_kl_subspace(z_x_mu, torch.ones(bs, device=x.device), subspace_dimensionality(node_mask, vae.n_dims))
    mean = torch.cat([z_x_mu, z_h_mu], dim=2)
    sigma = torch.cat([z_x_sigma.expand(-1, -1, 3), z_h_sigma], dim=2)
    if draws is not None and 'eps_enc' in draws:
        eps = draws['eps_enc'].to(x.device)
    else:
        eps = masked_noise(bs, n_nodes, vae.n_dims, vae.latent_node_nf, node_mask)
    z_xh = mean + sigma * eps
    x_rec, h_rec = vae.decoder._forward(z_xh, node_mask, edge_mask, context)
    rec = reconstruction_error(vae, torch.cat([x_rec, h_rec], dim=2), xh)
    loss = rec + vae.kl_weight * (kl_h + kl_x)
    return loss, {'loss_t': loss, 'rec_error': rec}


def latent_diffusion_nll(model, x, h, node_mask, edge_mask, context, draws=None):
    """EnLatentDiffusion.forward (:1136-1191): per-molecule loss (l2 in train() mode, NLL estimate in eval())."""
    bs, n_nodes = x.shape[0], x.shape[1]
    with torch.no_grad():                      # "always keep the encoder fixed" (:1155) -> fused inference kernels
        z_x_mu, _, z_h_mu, _ = model.vae.encode(x, h, node_mask, edge_mask, context)
        gamma_0 = model.gamma(torch.zeros(bs, 1, device=x.device))
        if draws is not None and 'eps_enc' in draws:
            eps = draws['eps_enc'].to(x.device)
        else:
            eps = masked_noise(bs, n_nodes, model.n_dims, model.vae.latent_node_nf, node_mask)
        z_xh = torch.cat([z_x_mu, z_h_mu], dim=2) + _sigma(gamma_0) * eps
    if model.trainable_ae:
        xh = torch.cat([x, h['categorical'], h['integer']], dim=2)
        x_rec, h_rec = model.vae.decoder._forward(z_xh, node_mask, edge_mask, context)
        loss_recon = reconstruction_error(model.vae, torch.cat([x_rec, h_rec], dim=2), xh)
    else:
        loss_recon = 0
    z_x, z_h = z_xh[:, :, :model.n_dims], z_xh[:, :, model.n_dims:]
    loss_ld, _ = diffusion_loss(model, z_x, z_h, node_mask, edge_mask, context, t0_always=not model.training,
                                draws=draws)
    if model.training and model.loss_type == 'l2':
        neg_log_const = torch.zeros_like(loss_ld)
    else:
        neg_log_const = -log_constants_h(model, node_mask)
    return loss_ld + loss_recon + neg_log_const


def compute_loss_and_nll(args, generative_model, nodes_dist, x, h, node_mask, edge_mask, context, draws=None):
    """qm9/losses.py:12-43: batch-mean of (loss - log p(N))."""
    bs, n_nodes, _ = x.shape
    if args.probabilistic_model != 'diffusion':
        raise ValueError(args.probabilistic_model)
    edge_mask = edge_mask.reshape(bs, n_nodes * n_nodes)
    # (host read; a captured step checks its static input once before capture: training.GraphedTrainStep)
    if not (x.is_cuda and torch.cuda.is_current_stream_capturing()) and float((x * (1 - node_mask)).abs().sum()) >= 1e-8:
        raise AssertionError("x is not masked")
    kw = {} if draws is None else {'draws': draws}
    nll = generative_model(x, h, node_mask, edge_mask, context, **kw)
    N = node_mask.squeeze(2).sum(1).long()
    log_pN = nodes_dist.log_prob(N)
    assert nll.shape == log_pN.shape
    nll = (nll - log_pN).mean(0)
    return nll, torch.zeros(1, device=nll.device), 0.
