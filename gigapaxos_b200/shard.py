"""Multi-GPU placement of Paxos groups.

Groups are independent replicated state machines, so the path shards with no data-path
collective: `home_gpu(paxosID) = |String.hashCode(paxosID)| mod n_gpus` -- the hash the
reference uses to spread coordinators (PaxosInstanceStateMachine.roundRobinCoordinator
:2251-2256).  *Packed* placement (the measured configuration) keeps all R replicas of a group
on its home GPU; *spread* placement puts replica j on GPU (home + j) mod n_gpus.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from .abi import java_string_hash


def home_gpu(paxos_id: str, n_gpus: int) -> int:
    h = java_string_hash(paxos_id)
    a = -h if h < 0 else h
    if a >= (1 << 31):  # Math.abs(Integer.MIN_VALUE) == Integer.MIN_VALUE
        a = 0
    return a % n_gpus


def replica_gpus(paxos_id: str, n_replicas: int, n_gpus: int, packed: bool = True) -> List[int]:
    h = home_gpu(paxos_id, n_gpus)
    if packed or n_gpus < n_replicas:
        return [h] * n_replicas
    return [(h + j) % n_gpus for j in range(n_replicas)]


def shard_names(names: Sequence[str], n_gpus: int) -> Dict[int, List[str]]:
    out: Dict[int, List[str]] = {r: [] for r in range(n_gpus)}
    for n in names:
        out[home_gpu(n, n_gpus)].append(n)
    return out


def shard_requests(names: Sequence[str], n_gpus: int) -> np.ndarray:
    """rank of every request's group (vectorised over a batch of paxos ids)."""
    return np.array([home_gpu(n, n_gpus) for n in names], dtype=np.int32)
