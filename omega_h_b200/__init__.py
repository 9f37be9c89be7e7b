"""omega_h_b200: B200-native refine path of omega_h behind the reference's own interface.
See DESIGN.md; the CUDA extension (lib/liboshb.so) is mandatory -- there is no CPU path."""
from ._lib import Lib, OshbError, default_lib  # noqa: F401
from .mesh import (EDGE, FACE, REGION, VERT, AdaptOpts, Mesh, adapt, build_box, last_pass_stats, refine_by_size,  # noqa: F401
                   simplex_degree)
from .osh_file import read_osh, write_osh  # noqa: F401
