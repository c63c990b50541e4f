"""Where the per-step cost of the global-time-step exchange goes: summary of the device-side globaltimer stamps that
`bench.py --trace PREFIX` leaves behind (one PREFIX.rank<r>.npy per rank, rows = timed steps, columns = FV_TRACE_* of
csrc/peer_mail.cuh).

    python scripts/exchange_attribution.py PREFIX [kernel_ms]

Only differences of stamps taken on the SAME device are used (globaltimer is per GPU and not synchronised across
devices).  For exchange s, published by the launch of step s-1 and consumed by the launch of step s:

    publish   = PUBLISHED(s) - LAST_WARP(s)          the last warp's stores into every peer's mailbox (incl. fences)
    wait      = WAIT_END(s) - WAIT_BEGIN(s)          how long the first warp of step s polled before all peers' values were there
    start-up  = WAIT_BEGIN(s) - KERNEL_BEGIN(s+1)    entry of the kernel -> first poll (ring requested, lane geometry)
    gap       = KERNEL_BEGIN(s+1) - PUBLISHED(s)     end of step s-1's last warp -> first warp of step s running (launch gap)
    blocking mode:  all_seen = ALL_SEEN(s) - PUBLISHED(s)   the last warp waiting inside step s-1's kernel
"""
import glob
import sys

import numpy as np

KB, WB, WE, LW, PB, AS = 0, 1, 2, 3, 4, 5


def us(x):
    x = np.asarray(x, dtype=np.float64) / 1e3
    return f"median {np.median(x):7.2f}  mean {x.mean():7.2f}  p90 {np.percentile(x, 90):7.2f}  max {x.max():7.2f} us"


def main():
    prefix = sys.argv[1]
    files = sorted(glob.glob(prefix + ".rank*.npy"), key=lambda f: int(f.rsplit("rank", 1)[1].split(".")[0]))
    if not files:
        raise SystemExit("no trace files " + prefix + ".rank*.npy")
    print(f"{len(files)} rank(s); stamps in ns of each device's own globaltimer")
    for f in files:
        t = np.load(f).astype(np.int64)
        rank = f.rsplit("rank", 1)[1].split(".")[0]
        t = t[1:-1]                                   # first / last row: neighbours outside the timed region
        pub = t[:, PB] - t[:, LW]
        print(f"rank {rank}: {len(t)} exchanges")
        print(f"   publish (last warp)            {us(pub)}")
        if (t[:, AS] > 0).all():
            print(f"   blocking wait in the kernel    {us(t[:, AS] - t[:, PB])}")
        if (t[:, WE] > 0).all():
            print(f"   consume wait (next launch)     {us(t[:, WE] - t[:, WB])}")
            # KERNEL_BEGIN is stored under the launch's own publish sequence: row s+1 holds the begin of the launch that consumes s
            kb_next = t[1:, KB]
            ok = kb_next > 0
            if ok.any():
                print(f"   kernel entry -> first poll     {us((t[:-1, WB] - kb_next)[ok])}")
                print(f"   published -> next kernel runs  {us((kb_next - t[:-1, PB])[ok])}")
                step = np.diff(t[:, KB][t[:, KB] > 0])
                if len(step):
                    print(f"   kernel entry to kernel entry   {us(step)}   (= the step time this rank saw)")


if __name__ == "__main__":
    main()
