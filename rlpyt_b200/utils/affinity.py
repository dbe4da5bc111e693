"""CPU placement for one-process-per-GPU runs: which host cores sit on the GPU's NUMA node.

Env workers write observations into the page-locked step buffer and the GPU's copy engine reads
them right away; on the 2-socket B200 hosts that DMA runs at ~6 GB/s when the lines were last
written by cores of the other socket and >40 GB/s when they are local (measured,
profiles/r01_sampler_probe.txt), so workers are pinned to the GPU-local cores.  (The reference
encodes placement in its "affinity code", rlpyt/utils/launching/affinity.py; only the
``workers_cpus`` / ``cuda_idx`` fields of that dict are consumed by the samplers.)
"""
import os


def _physical_index(cuda_idx):
    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    if visible:
        ids = [v.strip() for v in visible.split(",") if v.strip()]
        if cuda_idx < len(ids) and ids[cuda_idx].isdigit():
            return int(ids[cuda_idx])
    return cuda_idx


def _parse_ranges(text):
    cpus = []
    for part in text.split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        elif part.strip().isdigit():
            cpus.append(int(part))
    return cpus


def _from_nvml(phys, allowed):
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(phys)
    n_words = (max(allowed) + 64) // 64
    words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
    return [w * 64 + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]


def _from_topo(phys):
    """Parse the 'CPU Affinity' column of `nvidia-smi topo -m` (e.g. "32-63,96-127")."""
    import re
    import subprocess
    out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=10).stdout
    out = re.sub(r"\x1b\[[0-9;]*m", "", out)
    for line in out.splitlines():
        cols = line.split()
        if cols and cols[0] == f"GPU{phys}":
            for c in cols[1:]:
                if re.fullmatch(r"\d+(-\d+)?(,\d+(-\d+)?)*", c) and ("-" in c or "," in c):
                    return _parse_ranges(c)
    return []


def gpu_local_cpus(cuda_idx):
    """Cores local to GPU ``cuda_idx`` (NVML, else `nvidia-smi topo -m`), intersected with this
    process' allowed set; falls back to the allowed set."""
    allowed = sorted(os.sched_getaffinity(0))
    phys = _physical_index(cuda_idx)
    for probe in (lambda: _from_nvml(phys, allowed), lambda: _from_topo(phys)):
        try:
            local = [c for c in probe() if c in set(allowed)]
            if local and len(local) < len(allowed):
                return local
        except Exception as e:  # noqa: BLE001 - placement is best effort
            if os.environ.get("RLPYT_B200_DEBUG"):
                print("gpu_local_cpus probe failed:", repr(e))
    return allowed


def _sibling_groups(cpus):
    """Group hardware threads into physical cores (``thread_siblings_list`` of sysfs); each group is sorted
    and the groups are ordered by their first thread.  Falls back to one thread per core."""
    allowed, groups, seen = set(cpus), [], set()
    for c in sorted(cpus):
        if c in seen:
            continue
        sib = [c]
        try:
            with open(f"/sys/devices/system/cpu/cpu{c}/topology/thread_siblings_list") as f:
                sib = [x for x in _parse_ranges(f.read().strip()) if x in allowed] or [c]
        except OSError:
            pass
        sib = sorted(set(sib) | {c})
        seen.update(sib)
        groups.append(sib)
    return groups


def make_affinity(cuda_idx, n_workers=None, local_rank=0, ranks_per_node=1, node_share=8, smt_workers=False):
    """Affinity dict for the samplers.  The GPU-local NUMA node's PHYSICAL cores are the unit: ranks whose
    GPUs share a node split its cores evenly, and a rank never takes more than ``1/node_share`` of the
    whole machine's cores (an 8-GPU node: one eighth each), so a 1-GPU run uses the host share it would have
    in an 8-GPU run - weak scaling with fixed per-GPU resources.  The first core of the share belongs to
    the master; every worker gets its own core (``smt_workers``: its own hardware thread, two per core).
    Round 1 handed out hardware threads, not cores: at N >= 4 two ranks ended up on the two hyper-threads
    of the same cores and 14 workers were dealt onto 12 threads (SCALE_r01: sampling 68 -> 109 ms)."""
    local = gpu_local_cpus(cuda_idx)
    cores = _sibling_groups(local)
    all_cores = _sibling_groups(sorted(os.sched_getaffinity(0)))
    share = max(1, min(ranks_per_node, 8))
    same_node = [r for r in range(share) if gpu_local_cpus(r) == local] if share > 1 else [0]
    if len(same_node) > 1 and local_rank in same_node:
        k = same_node.index(local_rank)
        per = max(1, len(cores) // len(same_node))
        cores = cores[k * per:(k + 1) * per] or cores
    cap = max(2, len(all_cores) // max(1, node_share)) if node_share else len(cores)
    cores = cores[:cap]
    master = list(cores[0])
    pool = cores[1:] or cores
    slots = [[c[0]] for c in pool]
    if smt_workers:
        slots += [[c[1]] for c in pool if len(c) > 1]
    if not n_workers:
        n_workers = len(slots)
    n_workers = max(1, min(n_workers, len(slots)))
    workers_cpus = slots[:n_workers]
    return dict(cuda_idx=cuda_idx, master_cpus=master, workers_cpus=workers_cpus, set_affinity=True)
