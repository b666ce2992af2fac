"""Mnemonic histogram of the hot kernels from `cuobjdump -sass` of the built library (no GPU needed):
load / store widths, shared-memory atomics, barriers -- the SASS-level evidence for the kernel
choices (profiles/r2_sass_hot_kernels.md).

    python scripts/sass_summary.py > profiles/r2_sass_hot_kernels.md
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mh-spgemm_b200", "libmhb_spgemm.so")
WANT = {"k_num_compact_rowtwinsId": "k_num_compact_rowtwins<double>  (numeric, FEM-like rows: the headline kernel)",
        "k_num_hash_listIdLb0": "k_num_hash_list<double,false>  (numeric, claim-list hash, one block per row)",
        "k_num_hash_listIdLb1": "k_num_hash_list<double,true>   (numeric, claim-list hash, one warp per row)",
        "k_num_hash_blockId": "k_num_hash_block<double>  (numeric, 16 K-slot / global-pool rows)",
        "k_num_tinyId": "k_num_tiny<double>  (numeric, one thread per row)",
        "k_sym_bitmap_groupILi8": "k_sym_bitmap_group<8>  (symbolic, bitmap rows)",
        "k_sym_hash_groupILi32": "k_sym_hash_group<32>  (symbolic, tile hash, warp per row)",
        "k_mask_build": "k_mask_build  (family 1, one-pass mask builder with the chained scan)",
        "k_shard_pullId": "k_shard_pull<double>  (multi-GPU: one-sided halo pull over NVLink)"}
KEEP = ("LDG", "STG", "LDS", "STS", "ATOMS", "ATOMG", "ATOM", "RED", "SHFL", "REDUX", "MATCH", "UBLKCP", "LDGSTS", "POPC",
        "VOTE", "DFMA", "DADD", "DMUL", "BAR", "NANOSLEEP", "MEMBAR", "LD", "ST", "UTMALDG", "SYNCS")


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    print("# r2 — SASS of the hot kernels (`cuobjdump -sass mh-spgemm_b200/libmhb_spgemm.so`, sm_100a)\n")
    print("Counts of static SASS instructions per kernel (memory, atomic, shuffle, fp64 and barrier mnemonics only).\n"
          "What to read off: the headline numeric kernel has NO shared-memory atomic (`ATOMS`) at all -- its accumulator\n"
          "update is `LDS.64` / `DFMA` / `STS.64`; B's values are fetched with 64-bit `LDG.E.64.CONSTANT` (one 8-byte\n"
          "element per lane, a warp instruction covers two full 128-byte lines), its columns with 32-bit loads (one line);\n"
          "the hash kernels have exactly one `ATOMS.CAS` (the key claim) and the fp64 `ATOMS.CAST.SPIN.64` loop only on the\n"
          "hit path; the halo pull reads peer memory with plain `LDG.E(.64).STRONG.GPU` after one `LDG.E.64.STRONG.SYS` poll.\n"
          "No `UBLKCP` / `UTMALDG` (TMA) and no tensor-core instruction appears: B rows are 10-100 entries at arbitrary\n"
          "4-byte offsets (bulk copies need 16-byte alignment and size), and there is no dense contraction.\n")
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        key = [k for k in WANT if k in name]
        if not key:
            continue
        ops = collections.Counter()
        for m in re.finditer(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, re.M):
            if m.group(1).split(".")[0] in KEEP:
                ops[m.group(1)] += 1
        total = len(re.findall(r"^\s*/\*[0-9a-f]{4}\*/", f, re.M))
        print(f"### `{WANT[key[0]]}` — {total} SASS instructions\n")
        print("| " + " | ".join(f"{k} {v}" for k, v in sorted(ops.items())) + " |\n")


if __name__ == "__main__":
    main()
