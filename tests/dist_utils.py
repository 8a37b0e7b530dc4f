"""Helpers to run a function on N gloo/CPU processes inside a pytest test."""

from __future__ import annotations

import os
import socket
import traceback

import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _entry(rank: int, world_size: int, port: int, fn, args, errors):
    os.environ.update({"RANK": str(rank), "LOCAL_RANK": str(rank), "WORLD_SIZE": str(world_size),
                       "LOCAL_WORLD_SIZE": str(world_size), "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    try:
        import torch

        torch.set_num_threads(1)
        fn(rank, world_size, *args)
    except Exception:  # noqa: BLE001
        errors.put((rank, traceback.format_exc()))
        raise
    finally:
        import torch.distributed as dist

        if dist.is_initialized():
            dist.destroy_process_group()


def run_distributed(fn, world_size: int, *args) -> None:
    """Run ``fn(rank, world_size, *args)`` on ``world_size`` spawned CPU processes; re-raise the first failure."""
    ctx = mp.get_context("spawn")
    errors = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_entry, args=(r, world_size, port, fn, args, errors)) for r in range(world_size)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    failed = [p for p in procs if p.exitcode != 0]
    if failed:
        msgs = []
        while not errors.empty():
            msgs.append("rank %d:\n%s" % errors.get())
        for p in procs:
            if p.is_alive():
                p.kill()
        raise AssertionError("distributed test failed:\n" + "\n".join(msgs))
