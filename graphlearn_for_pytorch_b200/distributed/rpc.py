"""RPC control/data plane for multi-process deployments (worker mode across machines,
server-client mode).

Parity: reference python/distributed/rpc.py:56-529 -- init/shutdown, role-scoped all_gather
and barrier, callee registry, partition router.  It rides on torch.distributed.rpc
(TensorPipe, CPU tensors).  On a single NVSwitch box the *data* plane does not go through
here at all: sampling and feature kernels read peer HBM directly (parallel/); RPC remains
the bootstrap / control plane and the cross-machine fallback.
"""
import functools
import atexit
import collections
import logging
import threading
import time
from abc import ABC, abstractmethod
from typing import Callable, Dict, List, Optional

import torch
from torch.distributed import rpc

from .dist_context import DistContext, DistRole, get_context

_rpc_init_lock = threading.RLock()
_rpc_inited = False
_rpc_worker_names: Optional[Dict[DistRole, List[str]]] = None
_rpc_dynamic = False
_rpc_current_group_worker_names: Optional[List[str]] = None
_rpc_master_addr: Optional[str] = None
_rpc_master_port: Optional[int] = None

SERVER_INIT_CHECK_INTERVAL = 3.0
MAX_RETRY_TIMES = 60


def rpc_is_initialized() -> bool:
  return _rpc_inited


def _require_initialized(func):
  @functools.wraps(func)          # keeps the public signature (keyword names are part of the API)
  def wrapper(*args, **kwargs):
    if not rpc_is_initialized():
      raise RuntimeError('RPC has not been initialised; call init_rpc() first')
    return func(*args, **kwargs)
  wrapper.__name__ = func.__name__
  wrapper.__doc__ = func.__doc__
  return wrapper


@_require_initialized
def get_rpc_master_addr():
  return _rpc_master_addr


@_require_initialized
def get_rpc_master_port():
  return _rpc_master_port


@_require_initialized
def get_rpc_current_group_worker_names() -> List[str]:
  return _rpc_current_group_worker_names


@_require_initialized
def get_rpc_worker_names() -> Dict[DistRole, List[str]]:
  return _rpc_worker_names


# ----------------------------------------------------------------------------- role gather
class _GatherState(object):
  """Leader-side state of one role-scoped all_gather round."""

  def __init__(self, expected: int):
    self.expected = expected
    self.objects = {}
    self.future = torch.futures.Future()


_gather_lock = threading.Lock()
_gather_rounds: Dict[int, _GatherState] = {}
_gather_seq = 0


@rpc.functions.async_execution
def _role_gather_on_leader(seq: int, expected: int, worker_name: str, obj):
  """Runs on the group leader: completes once every member contributed."""
  with _gather_lock:
    st = _gather_rounds.get(seq)
    if st is None:
      st = _gather_rounds[seq] = _GatherState(expected)
    st.objects[worker_name] = obj
    done = len(st.objects) == st.expected
    if done:
      _gather_rounds.pop(seq, None)
  if done:
    st.future.set_result(dict(st.objects))
  return st.future


@_require_initialized
def all_gather(obj, timeout=None):
  """Gather `obj` from every process of the current role group -> {worker_name: obj}."""
  global _gather_seq
  names = _rpc_current_group_worker_names
  ctx = get_context()
  with _gather_lock:
    seq = _gather_seq
    _gather_seq += 1
  leader = names[0]
  _wait_for_member(leader)
  kw = {} if timeout is None else {'timeout': timeout}
  return rpc.rpc_sync(leader, _role_gather_on_leader, args=(seq, len(names), ctx.worker_name, obj), **kw)


@_require_initialized
def barrier(timeout=None):
  """Barrier over the current role group."""
  try:
    all_gather(None, timeout)
  except RuntimeError as e:  # pragma: no cover
    logging.error('failed to respond to barrier: %s', e)
    raise


@_require_initialized
def global_all_gather(obj, timeout=None):
  """Gather across *all* roles (servers + clients)."""
  t = rpc.api.UNSET_RPC_TIMEOUT if timeout is None else timeout
  return rpc.api._all_gather(obj, timeout=t)


@_require_initialized
def global_barrier(timeout=None):
  global_all_gather(None, timeout)


# ----------------------------------------------------------------------------- init / shutdown
def _own_worker_name():
  return rpc.get_worker_info().name


def _discover_name_by_rank(global_rank: int, fallback: str, timeout: float, interval: float = 0.2) -> str:
  """Name of the member that joined a dynamic RPC group with `global_rank`; `fallback` after `timeout` seconds
  (the conventional `<group>_<rank>` name is then resolved lazily by `_wait_for_member`)."""
  import time
  deadline = time.time() + min(float(timeout), 60.0)
  while True:
    try:
      return rpc.rpc_sync(global_rank, _own_worker_name, timeout=5.0)
    except Exception:  # noqa: BLE001  (not joined yet / transient connect error)
      if time.time() > deadline:
        return fallback
      time.sleep(interval)


def init_rpc(master_addr: str, master_port: int, num_rpc_threads: int = 16, rpc_timeout: float = 180,
             is_dynamic: bool = False):
  """Join the RPC world described by the current DistContext."""
  global _rpc_inited, _rpc_worker_names, _rpc_current_group_worker_names, _rpc_master_addr, _rpc_master_port
  with _rpc_init_lock:
    if _rpc_inited:
      return
    ctx = get_context()
    if ctx is None:
      raise RuntimeError("distributed context is not set; call init_worker_group / init_server / init_client")
    # TensorPipe channels for tensor payloads: multiplexed TCP + the basic fallback (the reference's choice,
    # rpc.py:256-258); GLT_B200_RPC_CHANNELS overrides the list, e.g. "cma,mpt_uv,basic" lets peers on ONE machine
    # move tensors with cross-memory attach (needs ptrace permission between the processes)
    import os
    channels = [c for c in os.environ.get('GLT_B200_RPC_CHANNELS', 'mpt_uv,basic').split(',') if c]
    opts = rpc.TensorPipeRpcBackendOptions(num_worker_threads=num_rpc_threads, rpc_timeout=rpc_timeout,
                                           init_method=f'tcp://{master_addr}:{master_port}',
                                           _transports=['uv'], _channels=channels)
    if is_dynamic:
      rpc.init_rpc(name=ctx.worker_name, rank=ctx.global_rank, world_size=None, rpc_backend_options=opts)
    else:
      rpc.init_rpc(name=ctx.worker_name, rank=ctx.global_rank, world_size=ctx.global_world_size,
                   rpc_backend_options=opts)
    _rpc_master_addr, _rpc_master_port = master_addr, master_port
    _rpc_inited = True
    global _rpc_dynamic
    _rpc_dynamic = bool(is_dynamic)
    if is_dynamic:
      # dynamic membership: derive names from the contexts instead of a global gather
      # (torch's dynamic RPC groups have no collectives; names follow the `<group>_<rank>` convention, the peer
      # role uses its default group name unless GLT_B200_PEER_GROUP names another one)
      import os
      from .dist_context import _DEFAULT_CLIENT_GROUP, _DEFAULT_SERVER_GROUP
      names = collections.defaultdict(list)
      names[ctx.role] = [f'{ctx.group_name}_{r}' for r in range(ctx.world_size)]
      if ctx.role == DistRole.CLIENT:
        g = os.environ.get('GLT_B200_PEER_GROUP')
        if g is not None:
          names[DistRole.SERVER] = [f'{g}_{r}' for r in range(ctx.num_servers())]
        else:
          # servers own global ranks [0, num_servers): ask each one for its name (it may use any group name), with
          # bounded retries while it is still joining (reference rpc.py:300-318)
          names[DistRole.SERVER] = [_discover_name_by_rank(r, f'{_DEFAULT_SERVER_GROUP}_{r}', rpc_timeout)
                                    for r in range(ctx.num_servers())]
      elif ctx.role == DistRole.SERVER:
        g = os.environ.get('GLT_B200_PEER_GROUP', _DEFAULT_CLIENT_GROUP)
        names[DistRole.CLIENT] = [f'{g}_{r}' for r in range(ctx.num_clients())]
      _rpc_worker_names = dict(names)
      _rpc_current_group_worker_names = names[ctx.role]
      return
    gathered = global_all_gather((ctx.role, ctx.world_size, ctx.rank))
    names = collections.defaultdict(dict)
    for name, (role, ws, rank) in gathered.items():
      names[role][rank] = name
    _rpc_worker_names = {role: [d[r] for r in sorted(d)] for role, d in names.items()}
    _rpc_current_group_worker_names = _rpc_worker_names[ctx.role]
    global_barrier()


def shutdown_rpc(graceful: bool = True):
  global _rpc_inited
  with _rpc_init_lock:
    if not _rpc_inited:
      return
    try:
      if graceful and not _rpc_dynamic:     # dynamic groups have no collectives: members just leave
        try:
          global_barrier()
        except Exception:  # noqa: BLE001
          pass
      rpc.shutdown(graceful=graceful)
    finally:
      _rpc_inited = False


atexit.register(shutdown_rpc, False)


# ----------------------------------------------------------------------------- partition router
class RpcDataPartitionRouter(object):
  """Round-robin over the workers that hold a given data partition."""

  def __init__(self, partition2workers: List[List[str]]):
    for pidx, workers in enumerate(partition2workers):
      if len(workers) == 0:
        raise ValueError(f'no RPC worker serves data partition {pidx}')
    self.partition2workers = partition2workers
    self._next = [0] * len(partition2workers)

  def get_to_worker(self, data_partition_idx: int) -> str:
    ws = self.partition2workers[data_partition_idx]
    i = self._next[data_partition_idx]
    self._next[data_partition_idx] = (i + 1) % len(ws)
    return ws[i]


@_require_initialized
def rpc_sync_data_partitions(num_data_partitions: int, current_partition_idx: int) -> List[List[str]]:
  """Tell everybody in the role group which partition this process serves."""
  ctx = get_context()
  gathered = all_gather((num_data_partitions, current_partition_idx))
  table = [[] for _ in range(num_data_partitions)]
  for name in _rpc_current_group_worker_names:
    n, p = gathered[name]
    if n != num_data_partitions:
      raise RuntimeError(f'{name} reports {n} data partitions, expected {num_data_partitions}')
    table[p].append(name)
  return table


# ----------------------------------------------------------------------------- callees
class RpcCalleeBase(ABC):
  """Server-side handler object; registered under an id that must match across processes."""

  @abstractmethod
  def call(self, *args, **kwargs):
    ...


_rpc_callee_lock = threading.RLock()
_rpc_callee_id = 0
_rpc_callee_pool: Dict[int, RpcCalleeBase] = {}


@_require_initialized
def rpc_register(callee: RpcCalleeBase) -> int:
  """Register a callee and check every process of the group got the same id."""
  global _rpc_callee_id
  with _rpc_callee_lock:
    callee_id = _rpc_callee_id
    _rpc_callee_id += 1
    _rpc_callee_pool[callee_id] = callee
  ids = all_gather(callee_id)
  for name, other in ids.items():
    if other != callee_id:
      raise RuntimeError(f'callee id mismatch: {other} on {name} vs {callee_id} here')
  return callee_id


def _rpc_call(callee_id, *args, **kwargs):
  return _rpc_callee_pool[callee_id].call(*args, **kwargs)


@_require_initialized
def rpc_request_async(worker_name: str, callee_id: int, args=None, kwargs=None):
  return rpc.rpc_async(worker_name, _rpc_call, args=(callee_id, *(args or ())), kwargs=kwargs)


@_require_initialized
def rpc_request(worker_name: str, callee_id: int, args=None, kwargs=None):
  return rpc_request_async(worker_name, callee_id, args, kwargs).wait()


_known_members = set()


def _wait_for_member(name: str, timeout: float = 120.0, interval: float = 0.2):
  """Dynamic membership has no rendezvous: a peer may simply not have joined yet.  Poll (bounded) until the
  RPC agent knows it -- the reference retries its dynamic joins the same way (rpc.py:286-318)."""
  if not _rpc_dynamic or name in _known_members:
    return
  import time
  deadline = time.time() + timeout
  while True:
    try:
      rpc.get_worker_info(name)
      _known_members.add(name)
      return
    except RuntimeError:
      if time.time() > deadline:
        raise
      time.sleep(interval)


@_require_initialized
def rpc_global_request_async(target_role: DistRole, role_rank: int, func: Callable, args=None, kwargs=None):
  """Call `func` on process `role_rank` of another role (e.g. client -> server)."""
  names = _rpc_worker_names.get(target_role)
  if names is None or role_rank >= len(names):
    ctx = get_context()
    prefix = {DistRole.SERVER: '_default_server', DistRole.CLIENT: '_default_client',
              DistRole.WORKER: '_default_worker'}[target_role]
    to = f'{prefix}_{role_rank}'
  else:
    to = names[role_rank]
  _wait_for_member(to)
  return rpc.rpc_async(to, func, args=args, kwargs=kwargs)


@_require_initialized
def rpc_global_request(target_role: DistRole, role_rank: int, func: Callable, args=None, kwargs=None):
  return rpc_global_request_async(target_role, role_rank, func, args, kwargs).wait()


@_require_initialized
def rpc_global_request_by_rank(global_rank: int, func: Callable, args=None, kwargs=None):
  return rpc.rpc_sync(global_rank, func, args=args, kwargs=kwargs)
