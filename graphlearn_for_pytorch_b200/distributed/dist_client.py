"""Client side of the server-client mode (parity: reference python/distributed/dist_client.py:24-101)."""
import logging
from typing import Optional

from .dist_context import DistRole, _set_client_context, get_context
from .dist_server import DistServer, _call_func_on_server
from .rpc import barrier, init_rpc, rpc_global_request, rpc_global_request_async, shutdown_rpc


def init_client(num_servers: int, num_clients: int, client_rank: int, master_addr: str, master_port: int,
                num_rpc_threads: int = 4, client_group_name: Optional[str] = None, is_dynamic: bool = False):
  """Declare this process as training client `client_rank` and join the RPC world."""
  _set_client_context(num_servers, num_clients, client_rank, client_group_name)
  init_rpc(master_addr, master_port, num_rpc_threads, is_dynamic=is_dynamic)


def shutdown_client():
  """Synchronise the clients, let client 0 stop every server, leave the RPC world."""
  ctx = get_context()
  if ctx is None:
    logging.warning('shutdown_client(): no distributed context')
    return
  if not ctx.is_client():
    raise RuntimeError('shutdown_client() must be called from a client process')
  barrier()
  if ctx.rank == 0:
    for srv in range(ctx.num_servers()):
      request_server(srv, DistServer.exit)
  shutdown_rpc()


def async_request_server(server_rank: int, func, *args, **kwargs):
  """Future of `func(server, *args)` evaluated on server `server_rank`."""
  return rpc_global_request_async(DistRole.SERVER, server_rank, _call_func_on_server, args=(func, *args),
                                  kwargs=kwargs)


def request_server(server_rank: int, func, *args, **kwargs):
  return rpc_global_request(DistRole.SERVER, server_rank, _call_func_on_server, args=(func, *args), kwargs=kwargs)
