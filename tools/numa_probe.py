"""Host topology + pinned-memory placement probe (one GPU is enough): prints the NUMA layout the box exposes and times
D2H / H2D copies of one GPU against page-locked buffers whose pages were first-touched under MPOL_BIND to each node,
MPOL_INTERLEAVE and the default policy.  Decides whether the proof / LDE host buffers should be placed by NUMA node.

    python tools/numa_probe.py [MiB]"""
import ctypes as C
import glob
import json
import os
import subprocess
import sys

import numpy as np
import torch

MPOL_DEFAULT, MPOL_BIND, MPOL_INTERLEAVE, SYS_set_mempolicy = 0, 2, 3, 238  # x86_64
libc = C.CDLL(None, use_errno=True)


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return f"<{e}>"


def set_policy(mode, nodes_mask, nnodes):
    if mode == MPOL_DEFAULT:
        return libc.syscall(SYS_set_mempolicy, MPOL_DEFAULT, None, C.c_ulong(0))
    mask = C.c_ulong(nodes_mask)
    return libc.syscall(SYS_set_mempolicy, mode, C.byref(mask), C.c_ulong(nnodes + 1))


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    nbytes = mib << 20
    nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
    print("nodes:", nodes)
    print(sh("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'"))
    print(sh("nvidia-smi topo -m | head -14"))
    print(sh("nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv"))
    for p in glob.glob("/sys/bus/pci/devices/*/vendor"):
        try:
            if open(p).read().strip() == "0x10de":
                d = os.path.dirname(p)
                cls = open(d + "/class").read().strip()
                if cls.startswith("0x0302") or cls.startswith("0x0300"):
                    print(os.path.basename(d), "numa_node", open(d + "/numa_node").read().strip())
        except OSError:
            pass
    print("affinity:", len(os.sched_getaffinity(0)), "cpus; current cpu", sh("cat /proc/self/stat | awk '{print $39}'"))
    torch.cuda.init()
    rt = torch.cuda.cudart()
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream()
    results = {}
    cases = [("default", MPOL_DEFAULT, 0)]
    if len(nodes) > 1:
        cases += [(f"bind{n}", MPOL_BIND, 1 << n) for n in nodes] + [("interleave", MPOL_INTERLEAVE, sum(1 << n for n in nodes))]
    for name, mode, mask in cases:
        rc = set_policy(mode, mask, max(nodes) + 1 if nodes else 1)
        a = np.empty(nbytes, dtype=np.uint8)
        a[::4096] = 1  # first touch under the policy
        set_policy(MPOL_DEFAULT, 0, 0)
        err = rt.cudaHostRegister(a.ctypes.data, nbytes, 1)
        t = torch.from_numpy(a)
        out = {"policy_rc": rc, "register": int(err)}
        for label, fn in (("d2h", lambda: t.copy_(dev, non_blocking=True)), ("h2d", lambda: dev.copy_(t, non_blocking=True))):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(5):
                fn()
            e1.record(s)
            torch.cuda.synchronize()
            out[label + "_gbs"] = round(5 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9, 2)
        rt.cudaHostUnregister(a.ctypes.data)
        results[name] = out
        print(name, out, flush=True)
        del t, a
    print(json.dumps(results))


if __name__ == "__main__":
    main()
