"""Loading of the golden fixtures (tests/golden/*.npz) and reconstruction of their injected streams."""
import glob
import json
import os

import numpy as np

import streams

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def env_cases():
    return sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "env_*.npz")))


def policy_cases():
    return sorted(os.path.basename(p)[7:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "policy_*.npz")))


class EnvCase:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, f"env_{name}.npz"))
        self.z = z
        self.meta = json.loads(str(z["meta"]))
        m = self.meta
        self.name, self.scenario, self.params = name, m["scenario"], m["params"]
        self.n, self.p, self.L, self.T, self.steps, self.seed = m["n"], m["p"], m["L"], m["T"], m["steps"], m["seed"]
        self.ge, self.planes, self.max_path_length = m["ge"], m["planes"], m["max_path_length"]
        self.actions = z["actions"]
        # same streams the generator injected into the reference (make_golden.py:run_env_case)
        self.chan_u = streams.uniforms_f32(self.seed * 16 + 3, (self.steps + 1, self.planes, self.n, self.n))
        self.cand = streams.prey_candidates(self.seed * 16 + 2, (self.steps, max(self.p, 1), 5))

    def unpack(self, key, s):
        """adjacency / channel masks of update s as dense uint8 (..., n, n)."""
        return np.unpackbits(self.z[key][s], axis=-1, count=self.n, bitorder="little")


class PolicyCase:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, f"policy_{name}.npz"))
        self.meta = json.loads(str(z["meta"]))
        self.n, self.D, self.B, self.L = (self.meta[k] for k in ("n", "D", "B", "L"))
        # constructor arguments beyond the defaults (tests/golden/make_golden_shapes.py: narrower layers, 'dot' attention)
        self.policy_kwargs = {k: (tuple(v) if isinstance(v, list) else v) for k, v in self.meta.get("policy_kwargs", {}).items()}
        self.weights = {k[3:]: z[k] for k in z.files if k.startswith("w::")}
        self.obs = z["obs"].reshape(self.B, self.n, self.D)
        self.avail = z["avail"].reshape(self.B, self.n, 5)
        self.adj = np.unpackbits(z["adj"], axis=-1, count=self.n, bitorder="little")
        self.chan = np.unpackbits(z["chan"], axis=-1, count=self.n, bitorder="little")
        self.probs, self.attn, self.logits = z["probs"], z["attn"], z["logits"]
