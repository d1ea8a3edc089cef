"""Pin a rank's host threads (and therefore its first-touch pinned staging memory) to the CPUs next to its GPU.

One process per GPU: result copies land in pinned host buffers at PCIe speed (37.8 MB per membrane position
at 2048^2); when the buffer lives on the other socket every byte also crosses the inter-socket link, which
8 ranks saturate.  The kernel exposes the GPU's local CPUs in sysfs; nothing here needs NVML.
"""
import os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def local_cpus(device_index):
    """CPUs of the NUMA node the GPU hangs off, or None when the platform does not say."""
    try:
        import torch
        props = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bus) as fh:
            cpus = _parse_cpulist(fh.read())
        return cpus or None
    except Exception:
        return None


def bind_to_gpu(device_index, ranks_on_node=1, local_rank=0):
    """Restrict this process to (a fair share of) the CPUs local to ``device_index``.  Returns the CPU set used,
    or None when nothing was changed."""
    cpus = local_cpus(device_index)
    if not cpus:
        return None
    try:
        allowed = os.sched_getaffinity(0)
    except AttributeError:
        return None
    cpus = sorted(cpus & allowed)
    if not cpus:
        return None
    try:
        os.sched_setaffinity(0, cpus)
    except OSError:
        return None
    return cpus
