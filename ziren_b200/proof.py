"""Decoder of the flat "ZKPF" shard-proof encoding (include/zkb200.h) into the fields of the
reference's `ShardProof` (crates/stark/src/types.rs:76-83)."""
from __future__ import annotations

import numpy as np

MAGIC = 0x46504B5A


class _R:
    def __init__(self, w):
        self.w, self.i = np.asarray(w, dtype=np.uint32), 0

    def u(self):
        v = int(self.w[self.i]); self.i += 1
        return v

    def take(self, n):
        v = self.w[self.i:self.i + n].copy(); self.i += n
        return v

    def s(self):
        n = self.u()
        raw = self.take((n + 3) // 4).tobytes()[:n]
        return raw.decode()


def parse(words) -> dict:
    r = _R(words)
    assert r.u() == MAGIC and r.u() == 1, "not a ZKPF proof"
    p = {"commitment": {"main_commit": r.take(8), "permutation_commit": r.take(8), "quotient_commit": r.take(8)}}
    chips = []
    for _ in range(r.u()):
        c = {"name": r.s(), "log_degree": r.u()}
        pw, mw, ew, nq = r.u(), r.u(), r.u(), r.u()
        for key, w in (("preprocessed", pw), ("main", mw), ("permutation", ew)):
            c[key] = {"local": r.take(4 * w).reshape(w, 4), "next": r.take(4 * w).reshape(w, 4)}
        c["quotient"] = r.take(16 * nq).reshape(nq, 4, 4)
        c["global_cumulative_sum"] = r.take(14)
        c["local_cumulative_sum"] = r.take(4)
        c["log_degree"] = c.pop("log_degree")            # field order of ChipOpenedValues (types.rs:51-64)
        chips.append(c)
    p["opened_values"] = {"chips": chips}
    p["chip_ordering"] = {c["name"]: i for i, c in enumerate(chips)}
    p["public_values"] = r.take(r.u())
    fri = {"commit_phase_commits": [r.take(8) for _ in range(r.u())]}
    fri["final_poly"], fri["pow_witness"] = r.take(4), r.u()
    qs = []
    for _ in range(r.u()):
        q = {"input_proof": [], "commit_phase_openings": []}
        for _ in range(r.u()):
            rows = [r.take(r.u()) for _ in range(r.u())]
            path = [r.take(8) for _ in range(r.u())]
            q["input_proof"].append({"opened_values": rows, "opening_proof": path})
        for _ in range(r.u()):
            sib = r.take(4)
            q["commit_phase_openings"].append({"sibling_value": sib, "opening_proof": [r.take(8) for _ in range(r.u())]})
        qs.append(q)
    fri["query_proofs"] = qs
    p["opening_proof"] = fri
    assert r.i == r.w.size, "trailing words in proof"
    # field order of ShardProof (crates/stark/src/types.rs:76-83)
    return {k: p[k] for k in ("commitment", "opened_values", "opening_proof", "chip_ordering", "public_values")}
