#!/usr/bin/env python
"""Regenerates the committed golden fixtures under tests/golden/ (run in the build container, where /root/reference exists).

For every world in WORLDS: synthesize the inputs (tools/synth.cpp, deterministic from the seed), run the UNMODIFIED reference
(oracle/_ref/arriba built from /root/reference/source + the htslib shim, deterministic list-node allocator on) and store
  fusions.tsv, fusions.discarded.tsv   the reference's own output files
  stages.npz                           fragment labels after the read-level cascade, candidate filters / read counts / e-values after
                                       every event-level stage (keyed by the 8-tuple candidate key), fragment-length statistics
The parity tests compare the product (CUDA on the GPU box, hostsim on CPU) and a fresh oracle run against these files."""
import json, os, shutil, sys, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import worldutil  # noqa: E402

WORLDS = {
    "tiny": dict(scale=0.001, genes=400, breakpoints=60, fragments=3000, read_length=101, seed=0xA881BA),
    "tiny_l151_shuffled": dict(scale=0.001, genes=300, breakpoints=40, fragments=2500, read_length=151, seed=11, extra=("--shuffle", "--varnames")),
    "tiny_mismapper_heavy": dict(scale=0.001, genes=300, breakpoints=60, fragments=4000, read_length=101, seed=5, extra=("--mismapper-frac", "0.3", "--paralog-frac", "0.15")),
}


def key_matrix(t):
    return np.stack([t["gene1"].astype(np.int64), t["gene2"], t["contig1"], t["contig2"], t["breakpoint1"], t["breakpoint2"], t["direction1"], t["direction2"]], axis=1)


def main():
    for name, params in WORLDS.items():
        out = os.path.join(HERE, name); os.makedirs(out, exist_ok=True)
        with tempfile.TemporaryDirectory() as d:
            p = dict(params); extra = p.pop("extra", ())
            prefix = os.path.join(d, "w")
            worldutil.run_synth(prefix, extra=extra, **p)
            worldutil.run_oracle(prefix, os.path.join(d, "oracle"))
            world = worldutil.World(prefix, os.path.join(d, "oracle"))
            for f in ("fusions.tsv", "fusions.discarded.tsv"):
                shutil.copy(os.path.join(d, "oracle", f), os.path.join(out, f))
            arrays = {"labels_after_read_filters": world.stage("rf_low_entropy")["frag_filter"], "labels_final": world.stage("ev_confidence")["frag_filter"],
                      "fragment_length": world.stage("fragment_length")["gap_mean_stddev_readlen"], "max_mate_gap": world.stage("find_fusions")["max_mate_gap"],
                      "candidate_keys": key_matrix(world.stage("find_fusions"))}
            seen = {}
            for stage, a in world.dumps:
                if not stage.startswith("ev_") and stage != "find_fusions":
                    continue
                occ = seen.get(stage, 0); seen[stage] = occ + 1
                tag = stage if occ == 0 else "%s_%d" % (stage, occ + 1)
                for col in ("filter", "split_reads1", "split_reads2", "discordant_mates", "evalue"):
                    arrays[tag + "." + col] = a[col]
            np.savez_compressed(os.path.join(out, "stages.npz"), **arrays)
            json.dump({"synth": {k: (list(v) if isinstance(v, tuple) else v) for k, v in params.items()}, "oracle_args": ["-f", "blacklist"], "det_alloc": True},
                      open(os.path.join(out, "params.json"), "w"), indent=1, sort_keys=True)
        print(name, "->", out, sorted(os.listdir(out)))


if __name__ == "__main__":
    main()
