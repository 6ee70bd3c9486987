"""Explainability data products: the val/test sweep that stores what the prototype head produced for every sample.

Mirror of ``load_data_and_model_products`` (reference: src/utils/explainability_utils.py:12-131) -- same arguments, same
file names, same two dictionaries and pickle schema, so ``explain.py`` of the reference keeps working:

    data_dict            inputs [N,3,(To),Ho,Wo], ys_gt [N], filenames (list)
    model_products_dict  fc_layer_weights [K,P], protoL_input_ [N,P,D], proto_dist_ [N,P],
                         occurrence_map_ [N,P,1,(T),H,W], ys_pred [N,classes]

What is different: the head runs in the CUDA library (``push_forward``), the softmax stays on the device, and results
leave the GPU through pinned staging buffers instead of one ``.cpu().numpy()`` + Python ``list.extend`` per tensor per
batch.  The f1 / confusion-matrix printout of the reference (``:88-114``) is a log line, not a data product; a compact
version of it is logged without the scikit-learn dependency.
"""
from __future__ import annotations

import os
import pickle
from typing import Callable, Dict, Tuple

import numpy as np
import torch


def _save_pickle(obj, path: str, log: Callable = print):
    with open(path, "wb") as handle:
        pickle.dump(obj, handle, protocol=pickle.HIGHEST_PROTOCOL)
    log(f"data successfully saved in {path}")


def _load_pickle(path: str, log: Callable = print):
    with open(path, "rb") as handle:
        obj = pickle.load(handle)
    log(f"data successfully loaded from {path}")
    return obj


def product_paths(mode: str, data_config: dict, root_dir_for_saving: str) -> Tuple[str, str]:
    """File names of the two pickles (reference: explainability_utils.py:18-28)."""
    filename = (
        f'{data_config["view"]}_'
        f'{data_config["frames"]}x{data_config["img_size"]}_'
        f'{data_config["interval_quant"]:.1f}x{data_config["interval_unit"]}_'
        f'{"all-Intervals" if data_config["iterate_intervals"] else ""}_'
        f"{mode}_data"
    )
    return (f'{data_config["dataset_root"]}/pickled_datasets/{filename}.pickle',
            f"{root_dir_for_saving}/{mode}/model_products.pickle")


@torch.no_grad()
def collect_model_products(model, dataloader, abstain_class: bool = True, keep_inputs: bool = True,
                           input_key: str = "cine", label_key: str = "target_AS", filename_key: str = "filename",
                           device=None) -> Tuple[Dict, Dict]:
    """One pass over ``dataloader`` (reference loop: explainability_utils.py:49-81).  Returns (data_dict,
    model_products_dict) with the reference's keys and dtypes (fp32 products, whatever dtype the loader yields for
    inputs/labels)."""
    if device is None:
        device = next(model.parameters()).device
    feats_l, dist_l, occ_l, pred_l, inputs_l, gt_l, filenames = [], [], [], [], [], [], []
    copy_stream = torch.cuda.Stream(device=device)
    pending = []     # (event, [(pinned, list)]) of batches whose device->host copies are in flight

    def drain(keep: int):
        while len(pending) > keep:
            ev, items = pending.pop(0)
            ev.synchronize()
            for host, sink in items:
                sink.append(host.numpy())

    for sample in dataloader:
        x = sample[input_key]
        y = sample[label_key]
        if keep_inputs:
            inputs_l.append(np.asarray(x.detach().cpu().numpy()))
        gt_l.append(np.asarray(y.detach().cpu().numpy() if torch.is_tensor(y) else y))
        names = sample.get(filename_key, []) if isinstance(sample, dict) else []
        filenames.extend(list(names))
        xb = x.to(device, non_blocking=True)
        feats, dist, occ, logits = model.push_forward(xb)
        if abstain_class:   # only the logits of the non-abstention classes enter the softmax (reference :66-70)
            prob = logits[:, : model.num_classes - 1].softmax(dim=1)
        else:
            prob = logits.softmax(dim=1)
        occ32 = occ.float()      # on the main stream: every tensor the copy stream reads is complete at `done`
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(device))
        items = []
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            for t, sink in ((feats, feats_l), (dist, dist_l), (occ32, occ_l), (prob, pred_l)):
                host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                host.copy_(t, non_blocking=True)
                t.record_stream(copy_stream)
                items.append((host, sink))
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        pending.append((ev, items))
        drain(keep=2)
    drain(keep=0)

    def cat(parts, empty_shape):
        return np.concatenate(parts, axis=0) if parts else np.zeros(empty_shape, dtype=np.float32)

    P, D = int(model.prototype_shape[0]), int(model.prototype_shape[1])
    data_dict = {
        "inputs": cat(inputs_l, (0,)) if keep_inputs else np.zeros((0,), dtype=np.float32),
        "ys_gt": np.concatenate([np.atleast_1d(g) for g in gt_l], axis=0) if gt_l else np.zeros((0,), dtype=np.int64),
        "filenames": filenames,
    }
    model_products_dict = {
        "fc_layer_weights": model.last_layer.weight.detach().cpu().numpy(),
        "protoL_input_": cat(feats_l, (0, P, D)),
        "proto_dist_": cat(dist_l, (0, P)),
        "occurrence_map_": cat(occ_l, (0, P, 1)),
        "ys_pred": cat(pred_l, (0, model.num_classes - (1 if abstain_class else 0))),
    }
    return data_dict, model_products_dict


def load_data_and_model_products(model, dataloader, mode, data_config, abstain_class, root_dir_for_saving, log=print):
    """Drop-in for the reference function of the same name (explainability_utils.py:12): loads the two pickles if both
    exist, else runs the sweep and writes them."""
    data_dict_path, model_products_path = product_paths(mode, data_config, root_dir_for_saving)
    os.makedirs(os.path.dirname(data_dict_path), exist_ok=True)
    os.makedirs(os.path.dirname(model_products_path), exist_ok=True)
    if os.path.exists(data_dict_path) and os.path.exists(model_products_path):
        data_dict = _load_pickle(data_dict_path, log)
        log(f"img  and labels and filenames of {mode}-dataset is loaded")
        model_products_dict = _load_pickle(model_products_path, log)
        log(f"model products for {mode}-dataset is loaded")
        return data_dict, model_products_dict
    log(f"model products not saved. running the epoch on {mode}-dataset to save the results.")
    data_dict, model_products_dict = collect_model_products(model, dataloader, abstain_class=abstain_class)
    ys_gt, ys_pred = data_dict["ys_gt"], model_products_dict["ys_pred"]
    if len(ys_gt):
        pred_class = ys_pred.argmax(axis=1)
        k = ys_pred.shape[1]
        f1 = []
        for c in range(k):
            tp = float(np.sum((pred_class == c) & (ys_gt == c)))
            fp = float(np.sum((pred_class == c) & (ys_gt != c)))
            fn = float(np.sum((pred_class != c) & (ys_gt == c)))
            f1.append(2 * tp / (2 * tp + fp + fn) if (2 * tp + fp + fn) > 0 else 0.0)
        log(f"f1 score is {np.asarray(f1)} with mean {float(np.mean(f1))}")
    _save_pickle(data_dict, data_dict_path, log)
    _save_pickle(model_products_dict, model_products_path, log)
    return data_dict, model_products_dict
