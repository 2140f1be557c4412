from .E_tracker import EssTracker
from .pnp_tracker import PnpTracker
