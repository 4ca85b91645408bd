"""horovodrun-shaped launcher (``python -m distributed_torch_horovod_gcp_b200.launch``)."""
from .run import run, main, parse_hosts, parse_hostfile, build_slots, Slot  # noqa: F401
