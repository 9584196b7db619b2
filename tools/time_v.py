"""`/usr/bin/time -v -o LOG cmd...` for boxes without GNU time: runs cmd, then writes the two lines the reference's
compare_container_benchmarks.py parses (parse_time_log: "Elapsed (wall clock) time", "Maximum resident set size") in GNU
time's own format.  Exit code = the command's."""
import os
import resource
import shutil
import subprocess
import sys
import time


def main():
    log, cmd = sys.argv[1], sys.argv[2:]
    gnu = shutil.which("time")
    if gnu and os.path.realpath(gnu) != os.path.realpath(sys.argv[0]):
        sys.exit(subprocess.call([gnu, "-v", "-o", log] + cmd))
    t0 = time.perf_counter()
    rc = subprocess.call(cmd)
    dt = time.perf_counter() - t0
    rss_kb = resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss          # kilobytes on Linux
    h, rem = divmod(dt, 3600.0)
    m, s = divmod(rem, 60.0)
    elapsed = f"{int(h)}:{int(m):02d}:{int(s):02d}" if h >= 1 else f"{int(m)}:{s:05.2f}"
    with open(log, "w") as f:
        f.write(f'\tCommand being timed: "{" ".join(cmd)}"\n')
        f.write(f"\tElapsed (wall clock) time (h:mm:ss or m:ss): {elapsed}\n")
        f.write(f"\tMaximum resident set size (kbytes): {rss_kb}\n")
        f.write(f"\tExit status: {rc}\n")
    sys.exit(rc)


if __name__ == "__main__":
    main()
