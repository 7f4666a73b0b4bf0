"""Host-side reference of the multi-shard protocol (tests only): record packing and the per-query merge with the same
semantics as the CUDA merge kernel (merge.cu): equal key strings collapse to their best score, then
(score desc, key length asc, first row asc), cut at `limit`."""
import numpy as np

RECORD_DTYPE = np.dtype([("scoreBits", "<u4"), ("keyLen", "<u4"), ("row", "<u8"), ("hash", "<u8")])
EMPTY_ROW = np.uint64(0xFFFFFFFFFFFFFFFF)


def shard_bounds(n_rows, world):
    return [n_rows * r // world for r in range(world + 1)]


def key_hash(key):
    """any deterministic 64-bit hash of the key string works for the host reference (FNV-1a here)."""
    h = 0xCBF29CE484222325
    for b in (key if isinstance(key, bytes) else key.encode("utf-32-le")):
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def pack_records(counts, ext, scores, keys, first_row, limit):
    """oracle results of one shard -> [nq, limit] records."""
    nq = len(counts)
    rec = np.zeros((nq, limit), RECORD_DTYPE)
    rec["row"] = EMPTY_ROW
    rec["hash"] = EMPTY_ROW
    for q in range(nq):
        for i in range(int(counts[q])):
            e = int(ext[q, i])
            rec[q, i] = (np.float32(scores[q, i]).view(np.uint32), len(keys[e]), first_row[e], key_hash(keys[e]))
    return rec


def merge_records(gathered, limit):
    """gathered: [shards, nq, limit] records -> ([nq, limit] records, counts)."""
    shards, nq, _ = gathered.shape
    out = np.zeros((nq, limit), RECORD_DTYPE)
    out["row"] = EMPTY_ROW
    out["hash"] = EMPTY_ROW
    counts = np.zeros(nq, np.uint32)
    for q in range(nq):
        best = {}
        for s in range(shards):
            for r in gathered[s, q]:
                if r["row"] == EMPTY_ROW:
                    continue
                score = np.uint32(r["scoreBits"]).view(np.float32)
                k = (int(r["hash"]), int(r["keyLen"]))
                cand = (-float(score), int(r["keyLen"]), int(r["row"]), r)
                if k not in best or cand[:3] < best[k][:3]:
                    best[k] = cand
        top = sorted(best.values(), key=lambda c: c[:3])[:limit]
        counts[q] = len(top)
        for i, c in enumerate(top):
            out[q, i] = c[3]
    return out, counts


# ---- global getSize / getLibSize over shards (SURVEY.md §8e "Global counters"): the protocol of stringsearchlib_b200/csrc/digest.cu
# in numpy — 16-byte digests per term / gram, distinct count per residue class, sum over the classes
def digests_of(units_2d_or_list):
    """16-byte digests (numpy void16) of unit sequences; any injective-enough function works for the protocol test."""
    import hashlib
    out = np.zeros(len(units_2d_or_list), dtype="V16")
    for i, u in enumerate(units_2d_or_list):
        out[i] = hashlib.blake2b(np.asarray(u, dtype=np.uint32).tobytes(), digest_size=16).digest()
    return out


def count_distinct_part(digests, part, n_parts):
    """distinct digests whose residue class is `part` (equal digests share a class)."""
    if len(digests) == 0:
        return 0
    lo = np.frombuffer(digests.tobytes(), dtype=np.uint64).reshape(-1, 2)[:, 0]
    sel = digests[(lo % np.uint64(n_parts)) == np.uint64(part)]
    return len(np.unique(sel))


def spec_terms(tables):
    off = tables["termOff"].astype(np.int64)
    return [tables["termChars"][off[i]:off[i + 1]] for i in range(len(off) - 1)]
