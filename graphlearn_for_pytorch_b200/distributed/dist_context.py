"""Distributed roles & context (parity: reference python/distributed/dist_context.py:20-212)."""
from enum import Enum
from typing import Dict, List, Optional


class DistRole(Enum):
  WORKER = 1   # worker mode: every process both samples and trains
  SERVER = 2   # server-client mode: sampling servers
  CLIENT = 3   # server-client mode: training clients


_DEFAULT_WORKER_GROUP = '_default_worker'
_DEFAULT_SERVER_GROUP = '_default_server'
_DEFAULT_CLIENT_GROUP = '_default_client'


class DistContext(object):
  """Position of this process: (role, group, rank in group) plus the global view."""

  def __init__(self, role: DistRole, group_name: str, world_size: int, rank: int,
               global_world_size: int, global_rank: int):
    assert 0 <= rank < world_size and 0 <= global_rank < global_world_size
    self.role, self.group_name = role, group_name
    self.world_size, self.rank = world_size, rank
    self.global_world_size, self.global_rank = global_world_size, global_rank

  def __repr__(self):
    return (f'DistContext(role={self.role}, group={self.group_name}, rank={self.rank}/{self.world_size}, '
            f'global={self.global_rank}/{self.global_world_size})')

  def __eq__(self, o):
    return isinstance(o, DistContext) and vars(self) == vars(o)

  def is_worker(self):
    return self.role == DistRole.WORKER

  def is_server(self):
    return self.role == DistRole.SERVER

  def is_client(self):
    return self.role == DistRole.CLIENT

  def num_servers(self) -> int:
    if self.role == DistRole.SERVER:
      return self.world_size
    if self.role == DistRole.CLIENT:
      return self.global_world_size - self.world_size
    return 0

  def num_clients(self) -> int:
    if self.role == DistRole.CLIENT:
      return self.world_size
    if self.role == DistRole.SERVER:
      return self.global_world_size - self.world_size
    return 0

  @property
  def worker_name(self) -> str:
    return f'{self.group_name}_{self.rank}'


_dist_context: Optional[DistContext] = None
_clients_to_servers: Optional[Dict[int, List[int]]] = None


def get_context() -> Optional[DistContext]:
  return _dist_context


def get_clients_to_servers():
  return _clients_to_servers


def _set_context(ctx: Optional[DistContext]):
  global _dist_context
  _dist_context = ctx


def _set_worker_context(world_size: int, rank: int, group_name: Optional[str] = None):
  _set_context(DistContext(DistRole.WORKER, group_name or _DEFAULT_WORKER_GROUP, world_size, rank, world_size, rank))


def _set_server_context(num_servers: int, server_rank: int, server_group_name: Optional[str] = None,
                        num_clients: int = 0):
  assert num_servers > 0
  _set_context(DistContext(DistRole.SERVER, server_group_name or _DEFAULT_SERVER_GROUP, num_servers, server_rank,
                           num_servers + num_clients, server_rank))


def _set_client_context(num_servers: int, num_clients: int, client_rank: int,
                        client_group_name: Optional[str] = None):
  assert num_servers > 0 and num_clients > 0
  _set_context(DistContext(DistRole.CLIENT, client_group_name or _DEFAULT_CLIENT_GROUP, num_clients, client_rank,
                           num_servers + num_clients, num_servers + client_rank))
  assign_server_by_order()


def assign_server_by_order():
  """Client c talks to servers {s : s % num_clients == c} when there are at least as many
  servers as clients, otherwise to server c % num_servers (round-robin both ways)."""
  global _clients_to_servers
  ctx = get_context()
  assert ctx is not None and ctx.is_client()
  ns, nc = ctx.num_servers(), ctx.num_clients()
  table = {}
  for c in range(nc):
    if ns >= nc:
      table[c] = [s for s in range(ns) if s % nc == c]
    else:
      table[c] = [c % ns]
  _clients_to_servers = table
  return table[ctx.rank]


def init_worker_group(world_size: int, rank: int, group_name: Optional[str] = None):
  """Worker-mode entry point: declare this process as worker `rank` of `world_size`."""
  _set_worker_context(world_size, rank, group_name)
