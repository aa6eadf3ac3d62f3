#!/usr/bin/env python
"""Generate the committed golden fixtures from the reference's shipped RLlib checkpoints.

Run HERE (container with /root/reference mounted), never on the GPU box:

    python tests/golden/make_golden.py

Reads   /root/reference/Results/**/checkpoint_1250/checkpoint-1250   (two nested pickles; the
        only non-numpy classes are ray.rllib.utils.filter.{MeanStdFilter,RunningStat} which are
        stubbed)
Writes  tests/golden/ckpt_<Arch>[_TVel].npz   seed-0 trial of each published architecture:
            names            policy ids (order of the checkpoint's "state" dict)
            <pid>/theta      flat FP32 parameter vector in checkpoint variable order
                             fc_1(k,b) fc_value_1(k,b) fc_2(k,b) fc_value_2(k,b) fc_out(k,b) value_out(k,b)
            <pid>/shapes     int array [12,2] of the variable shapes (bias -> [n,0])
            <pid>/filter_n   int64 count;  <pid>/filter_M, <pid>/filter_S  float64 [D]
            <pid>/adam_m, adam_v, beta_powers   (FullyDecentral only: pins TF1 Adam slot layout)
        tests/golden/learner_stats.json   final learner stats of all 360 policies (loss identity KAT)
        tests/golden/param_counts.json    (num_inputs, num_hidden, num_out, weights) rows of
                                          Results/experiment_2_nn_hidden_sizes_comparison.csv
"""
import glob
import io
import json
import os
import pickle

import numpy as np

REF = "/root/reference/Results"
OUT = os.path.dirname(os.path.abspath(__file__))

VAR_ORDER = ["fc_1", "fc_value_1", "fc_2", "fc_value_2", "fc_out", "value_out"]


class _Stub:
    def __setstate__(self, s):
        self.__dict__.update(s)


class StubUnpickler(pickle.Unpickler):
    def find_class(self, mod, name):
        try:
            return super().find_class(mod, name)
        except Exception:
            return type(name, (_Stub,), {})


def load_ckpt(path):
    top = StubUnpickler(open(path, "rb")).load()
    worker = StubUnpickler(io.BytesIO(top["worker"])).load()
    return top, worker


def flat_vars(pid, od, suffix=""):
    """suffix '' -> weights; '/Adam' or '/Adam_1' -> slots (looked up in _optimizer_variables)."""
    src = od if not suffix else od["_optimizer_variables"]
    chunks, shapes = [], []
    for layer in VAR_ORDER:
        for part in ("kernel", "bias"):
            key = f"{pid}/{layer}/{part}"
            if suffix:
                key = f"{pid}/{key}{suffix}"
            a = np.asarray(src[key], dtype=np.float32)
            shapes.append(list(a.shape) + [0] * (2 - a.ndim))
            chunks.append(a.reshape(-1))
    return np.concatenate(chunks), np.asarray(shapes, dtype=np.int64)


def main():
    exp1 = sorted(glob.glob(f"{REF}/experiment_1_models_architectures_on_flat/HF_10_QuantrupedMultiEnv_*"))
    exp3 = sorted(glob.glob(f"{REF}/experiment_3_models_curriculum_tvel/Tvel_QuantrupedMultiEnv_*"))
    stats = {}
    for arch_dir in exp1 + exp3:
        arch = os.path.basename(arch_dir).split("QuantrupedMultiEnv_")[1]
        trials = sorted(glob.glob(f"{arch_dir}/PPO_*"))
        for ti, trial in enumerate(trials):
            ck = f"{trial}/checkpoint_1250/checkpoint-1250"
            top, worker = load_ckpt(ck)
            meta = pickle.load(open(ck + ".tune_metadata", "rb"))
            learner = top["train_exec_impl"]["info"]["learner"]
            stats[f"{arch}/{ti}"] = {
                "ray_version": meta.get("ray_version"),
                "time_total": float(meta.get("time_total")),
                "policies": {
                    pid: {k: float(v) for k, v in d.items() if k != "model"} for pid, d in learner.items()
                },
            }
            if ti != 0:
                continue
            out = {"names": np.array(list(worker["state"].keys()))}
            for pid, od in worker["state"].items():
                theta, shapes = flat_vars(pid, od)
                out[f"{pid}/theta"] = theta
                out[f"{pid}/shapes"] = shapes
                flt = worker["filters"][pid]
                assert flt.clip is None and flt.demean and flt.destd and not len(flt.buffer.__dict__.get("_M", [0])) == 0
                out[f"{pid}/filter_n"] = np.int64(flt.rs._n)
                out[f"{pid}/filter_M"] = np.asarray(flt.rs._M, dtype=np.float64)
                out[f"{pid}/filter_S"] = np.asarray(flt.rs._S, dtype=np.float64)
                if arch == "FullyDecentral":
                    out[f"{pid}/adam_m"], _ = flat_vars(pid, od, "/Adam")
                    out[f"{pid}/adam_v"], _ = flat_vars(pid, od, "/Adam_1")
                    ov = od["_optimizer_variables"]
                    out[f"{pid}/beta_powers"] = np.array(
                        [ov[f"{pid}/beta1_power"], ov[f"{pid}/beta2_power"]], dtype=np.float32
                    )
            np.savez_compressed(f"{OUT}/ckpt_{arch}.npz", **out)
            print(arch, {k: v.shape for k, v in out.items() if k.endswith("theta")})
    json.dump(stats, open(f"{OUT}/learner_stats.json", "w"), indent=0, sort_keys=True)

    import csv

    rows = set()
    with open(f"{REF}/experiment_2_nn_hidden_sizes_comparison.csv", encoding="utf-8-sig") as f:
        for r in csv.DictReader(f):
            rows.add((r["approach"], int(r["num_inputs"]), int(r["num_hidden"]), int(r["num_out"]),
                      int(r["num_contr"]), int(r["weights"])))
    json.dump(sorted(rows), open(f"{OUT}/param_counts.json", "w"))
    print(len(stats), "trials;", len(rows), "param-count rows")


if __name__ == "__main__":
    main()
