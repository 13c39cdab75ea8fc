# same star-exports as the reference's src/models/__init__.py:1-2 (mspn is out of scope)
from .hourglass import *  # noqa: F401,F403
