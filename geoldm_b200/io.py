"""On-disk formats of the reference: experiment directories and xyz sample files.

* ``load_checkpoint`` reads what main_qm9.py / main_geom_drugs.py write and eval_analyze.py:127-157 reads:
  ``<dir>/args.pickle`` (pickled argparse.Namespace) and ``<dir>/generative_model[_ema].npy`` (``torch.save`` of the
  state_dict, utils.py:20-27), builds the CUDA model through ``get_latent_diffusion`` and loads the weights strictly.
* ``save_xyz_file`` / ``load_molecule_xyz`` follow qm9/visualizer.py:18-58 byte for byte ("<n>\\n\\n" header, one
  "El x y z" line per atom with 9 decimals, files ``<name>_%03d.txt``).
* ``analyze_and_save`` is eval_analyze.py:35-67 on top of the CUDA sampler and the device stability metric.
"""
from __future__ import annotations

import os
import pickle
import time
from os.path import join

import torch

from .models import get_latent_diffusion
from .sampling import sample
from .stability import analyze_stability_for_molecules


def save_model(model, path):
    torch.save(model.state_dict(), path)


def load_model(model, path):
    model.load_state_dict(torch.load(path, map_location="cpu"))
    model.eval()
    return model


def load_checkpoint(model_path, device, dataset_info, mma_mode="auto", use_ema=None):
    """-> (generative_model, nodes_dist, prop_dist, args).  ``dataset_info`` is the reference's
    configs/datasets_config.py entry for args.dataset (atom_decoder, n_nodes histogram, max_n_nodes).
    Both files are Python pickles (the reference's format): load experiment directories from trusted sources only."""
    with open(join(model_path, 'args.pickle'), 'rb') as f:
        args = pickle.load(f)
    if not hasattr(args, 'normalization_factor'):
        args.normalization_factor = 1
    if not hasattr(args, 'aggregation_method'):
        args.aggregation_method = 'sum'
    args.mma_mode = mma_mode
    if len(getattr(args, 'conditioning', [])) > 0:
        # the property prior needs the QM9 training set (qm9/models.py:60-61); keep the context width, drop the prior
        args.context_node_nf = getattr(args, 'context_node_nf', len(args.conditioning))
        args.conditioning = []
    if getattr(args, 'ae_path', None) is not None:
        args.ae_path = None                     # first-stage weights are part of the state_dict ('vae.*')
    model, nodes_dist, prop_dist = get_latent_diffusion(args, device, dataset_info, None)
    ema = (getattr(args, 'ema_decay', 0) > 0) if use_ema is None else use_ema
    fn = 'generative_model_ema.npy' if ema else 'generative_model.npy'
    state = torch.load(join(model_path, fn), map_location=device)
    model.load_state_dict(state)
    return model.eval(), nodes_dist, prop_dist, args


def save_xyz_file(path, one_hot, charges, positions, dataset_info, id_from=0, name='molecule', node_mask=None):
    os.makedirs(path, exist_ok=True)
    bs, n_max = one_hot.shape[0], one_hot.shape[1]
    if node_mask is not None:
        counts = node_mask.reshape(bs, -1).sum(1).cpu()
    else:
        counts = torch.full((bs,), n_max)
    atoms = torch.argmax(one_hot, dim=2).cpu()
    pos = positions.detach().cpu()
    decoder = dataset_info['atom_decoder']
    for b in range(bs):
        n = int(counts[b])
        lines = ["%d\n\n" % counts[b]]
        for a in range(n):
            lines.append("%s %.9f %.9f %.9f\n" % (decoder[int(atoms[b, a])], pos[b, a, 0], pos[b, a, 1], pos[b, a, 2]))
        with open(path + name + '_' + "%03d.txt" % (b + id_from), "w") as f:
            f.write("".join(lines))


def load_molecule_xyz(file, dataset_info):
    with open(file, encoding='utf8') as f:
        n_atoms = int(f.readline())
        f.readline()
        rows = f.readlines()
    one_hot = torch.zeros(n_atoms, len(dataset_info['atom_decoder']))
    charges = torch.zeros(n_atoms, 1)
    positions = torch.zeros(n_atoms, 3)
    for i in range(n_atoms):
        parts = rows[i].split(' ')
        one_hot[i, dataset_info['atom_encoder'][parts[0]]] = 1
        positions[i] = torch.tensor([float(v) for v in parts[1:]])
    return positions, one_hot, charges


def analyze_and_save(args, eval_args, device, generative_model, nodes_dist, prop_dist, dataset_info, n_samples=10,
                     batch_size=10, save_to_xyz=False, **sampler_kwargs):
    batch_size = min(batch_size, n_samples)
    assert n_samples % batch_size == 0
    molecules = {'one_hot': [], 'x': [], 'node_mask': []}
    start = time.time()
    for i in range(n_samples // batch_size):
        nodesxsample = nodes_dist.sample(batch_size)
        one_hot, charges, x, node_mask = sample(args, device, generative_model, dataset_info, prop_dist=prop_dist,
                                                nodesxsample=nodesxsample, **sampler_kwargs)
        molecules['one_hot'].append(one_hot)
        molecules['x'].append(x)
        molecules['node_mask'].append(node_mask)
        done = (i + 1) * batch_size
        print('\t %d/%d Molecules generated at %.4f secs/sample' % (done, n_samples, (time.time() - start) / done))
        if save_to_xyz:
            save_xyz_file(join(eval_args.model_path, 'eval/analyzed_molecules/'), one_hot, charges, x, dataset_info,
                          i * batch_size, name='molecule', node_mask=node_mask)
    molecules = {k: torch.cat(v, dim=0) for k, v in molecules.items()}
    return analyze_stability_for_molecules(molecules, dataset_info)
