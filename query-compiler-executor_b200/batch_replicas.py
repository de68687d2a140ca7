"""Query-level parallelism for batches of independent queries (SURVEY.md 8e, config 5:
"round-robin whole small queries across GPUs, buffer result lines, emit in input order").

Replicas only: every rank holds the relations it needs and runs whole queries through the
unsharded host layer (execute_queries' loop, src/utilities.c:289-300, one query at a time);
there is no data-path collective.  Query i goes to rank i mod world; rank 0 collects the
output blocks (the count lines of stacked filters, src/filter.c:32, travel with their query's
result line) and emits them in input order, which is the reference's stdout order.
Large joins use the sharded executor (shardexec.py) instead.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence


def split_queries(text: str) -> List[str]:
    """Query lines of a batch; `F` lines (batch separators, src/parsing.c) and blanks are skipped."""
    return [ln for ln in (x.strip() for x in text.splitlines()) if ln and ln != "F"]


def run_batch_replicated(run_one: Callable[[str], str], queries: Sequence[str], dist, rank: int,
                         world: int) -> Optional[str]:
    """run_one(query_line) -> the bytes the reference prints for that query.  Returns the
    whole batch's stdout on rank 0, None elsewhere."""
    mine = [(i, run_one(q)) for i, q in enumerate(queries) if i % world == rank]
    if world == 1:
        return "".join(out for _, out in mine)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    if rank != 0:
        return None
    blocks = [""] * len(queries)
    for part in gathered:
        for i, out in part:
            blocks[i] = out
    return "".join(blocks)


def host_layer_runner(host_lib) -> Callable[[str], str]:
    """run_one on top of libqce_host.so's in-process batch entry (relations already resident)."""
    import ctypes as C

    def run_one(q: str) -> str:
        buf = C.create_string_buffer(1 << 16)
        failed = C.c_int(0)
        n = host_lib.qce_host_run_batch((q + "\n").encode(), buf, 1 << 16, C.byref(failed))
        if n < 0:
            raise RuntimeError("host layer could not parse the query")
        return buf.value.decode()
    return run_one
