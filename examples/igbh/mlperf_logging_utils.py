"""MLPerf-style event logging for the IGBH jobs (counterpart of the reference's
examples/igbh/mlperf_logging_utils.py): thin wrapper over utils.tracing.EventLogger which prints
':::MLLOG {json}' lines (INIT_START/STOP, RUN_START/STOP, EPOCH_*, EVAL_*, EVAL_ACCURACY, hyper-parameters)."""
import os.path as osp
import sys

sys.path.insert(0, osp.dirname(osp.dirname(osp.dirname(osp.abspath(__file__)))))
from graphlearn_for_pytorch_b200.utils import EventLogger  # noqa: E402


def get_mlperf_logger(path=None, rank: int = 0) -> EventLogger:
  return EventLogger(path, rank)


def submission_info(log: EventLogger, benchmark: str = 'gnn', submitter: str = 'reference', platform: str = 'B200'):
  for k, v in (('SUBMISSION_BENCHMARK', benchmark), ('SUBMISSION_ORG', submitter), ('SUBMISSION_DIVISION', 'closed'),
               ('SUBMISSION_STATUS', 'onprem'), ('SUBMISSION_PLATFORM', platform)):
    log.event(k, v)
