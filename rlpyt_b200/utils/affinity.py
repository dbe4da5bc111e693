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


def make_affinity(cuda_idx, n_workers, local_rank=0, ranks_per_node=1, reserve_master=1):
    """Affinity dict for the samplers: this rank's share of the GPU-local cores, one core reserved
    for the master, the rest dealt round-robin to ``n_workers`` workers."""
    local = gpu_local_cpus(cuda_idx)
    # ranks whose GPUs share a NUMA node split its cores evenly (by local rank order)
    share = max(1, min(ranks_per_node, 8))
    same_node = [r for r in range(share) if gpu_local_cpus(r) == local] if share > 1 else [0]
    if len(same_node) > 1 and local_rank in same_node:
        k = same_node.index(local_rank)
        per = max(1, len(local) // len(same_node))
        local = local[k * per:(k + 1) * per] or local
    master = local[:reserve_master]
    pool = local[reserve_master:] or local
    n_workers = max(1, n_workers)
    workers_cpus = [[pool[i % len(pool)]] for i in range(n_workers)]
    return dict(cuda_idx=cuda_idx, master_cpus=master, workers_cpus=workers_cpus, set_affinity=True)
