"""How many CPU threads may this process really use?  (test / bench infrastructure)

`os.cpu_count()` reports the machine's logical CPUs; inside a container the usable share is bounded by the
scheduler affinity mask and by the cgroup CPU quota.  Asking torch for more threads than that oversubscribes
the cores and makes the CPU oracle 10-20x slower, so every CPU-timed leg sizes its thread pool with this."""
import os


def usable_cpus() -> int:
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if q > 0:
                    n = min(n, max(1, int(q / p + 0.5)))
        except Exception:
            pass
    return max(1, n)


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"
