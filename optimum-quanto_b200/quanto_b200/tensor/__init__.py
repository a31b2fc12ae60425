from .activations import *  # noqa: F401,F403
from .core import *  # noqa: F401,F403
from .grouped import *  # noqa: F401,F403
from .optimizers import *  # noqa: F401,F403
from .packed import *  # noqa: F401,F403
from .qtensor import *  # noqa: F401,F403
from .qtype import *  # noqa: F401,F403
from .weights import *  # noqa: F401,F403
