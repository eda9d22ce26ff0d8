"""Optional worker processes for the embarrassingly parallel host loops of the DCT-weight path (per-block view probes,
per-cell L-BFGS-B exponent fits: 1.2 s + 2.8 s of single-threaded Python for 8 views on a 512x1024x1024 frame).

``MVF_HOST_WORKERS=N`` (N >= 2; default: this rank's share of the host cores minus one, at most 16) starts N plain subprocesses ``python -m mvregfus_b200._host_workers`` that
import numpy/scipy only (no torch, nothing of the caller's process is forked or re-imported) and exchange
length-prefixed pickles over pipes.  With 0 or 1 the loops run serially in this process.  Any worker
failure or timeout kills the workers and reruns the tasks serially - per item the code is the same function, so the
results are identical bit for bit either way."""
import atexit
import os
import pickle
import struct
import subprocess
import sys
import threading

_procs = []


def workers():
    """MVF_HOST_WORKERS=N; unset: the host cores of this rank (cores / ranks on the node, at most 16) minus one."""
    try:
        return max(0, int(os.environ["MVF_HOST_WORKERS"]))
    except (KeyError, ValueError):
        pass
    try:
        share = len(os.sched_getaffinity(0)) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    except (AttributeError, ValueError):
        share = 1
    n = min(16, share - 1)
    return n if n >= 2 else 0


def shutdown():
    global _procs
    for p in _procs:
        try:
            p.kill()
        except Exception:
            pass
    _procs = []


atexit.register(shutdown)


def _start(n):
    global _procs
    if len(_procs) == n and all(p.poll() is None for p in _procs):
        return _procs
    shutdown()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    env["OMP_NUM_THREADS"] = env["OPENBLAS_NUM_THREADS"] = env["MKL_NUM_THREADS"] = "1"
    _procs = [subprocess.Popen([sys.executable, "-m", "mvregfus_b200._host_workers"], stdin=subprocess.PIPE,
                               stdout=subprocess.PIPE, env=env, cwd=root) for _ in range(n)]
    return _procs


def _send(p, obj):
    data = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    p.stdin.write(struct.pack("<Q", len(data)) + data)
    p.stdin.flush()


def _recv(p):
    head = p.stdout.read(8)
    if len(head) != 8:
        raise EOFError("worker closed its pipe")
    (n,) = struct.unpack("<Q", head)
    data = p.stdout.read(n)
    if len(data) != n:
        raise EOFError("short read from worker")
    ok, val = pickle.loads(data)
    if not ok:
        raise RuntimeError(val)
    return val


def pmap(func, tasks, min_tasks=2, timeout=600):
    """[func(t) for t in tasks]; with MVF_HOST_WORKERS >= 2 the tasks are spread over worker subprocesses."""
    tasks = list(tasks)
    n = min(workers(), len(tasks))
    if n <= 1 or len(tasks) < min_tasks:
        return [func(t) for t in tasks]
    results, errors = [None] * len(tasks), []
    try:
        procs = _start(workers())[:n]

        def drive(w):
            try:
                for i in range(w, len(tasks), n):
                    _send(procs[w], (func.__name__, tasks[i]))
                    results[i] = _recv(procs[w])
            except Exception as exc:  # noqa: BLE001
                errors.append(exc)

        threads = [threading.Thread(target=drive, args=(w,), daemon=True) for w in range(n)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout)
        if errors or any(t.is_alive() for t in threads):
            raise RuntimeError(errors[0] if errors else "host workers timed out")
        return results
    except Exception:  # noqa: BLE001
        shutdown()
        return [func(t) for t in tasks]


def chunks(n_items, per_worker=4):
    """Index ranges [(a, b), ...] covering range(n_items): ~per_worker tasks per worker (one range when serial)."""
    k = max(1, min(n_items, max(1, workers()) * per_worker))
    step = -(-n_items // k)
    return [(a, min(n_items, a + step)) for a in range(0, n_items, step)]
