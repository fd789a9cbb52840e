"""Process-group bring-up (mirror of the reference's ``src/utils/distributed.py:20-47``)."""
from avjepa_b200.dist import init_distributed  # noqa: F401
