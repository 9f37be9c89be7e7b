"""omega_h_b200: B200-native refine path of omega_h behind the reference's own interface.
See DESIGN.md; the CUDA extension (lib/liboshb.so) is mandatory -- there is no CPU path."""
from ._lib import Lib, OshbError, default_lib  # noqa: F401
from .mesh import (EDGE, FACE, REGION, VERT, AdaptOpts, Mesh, adapt, build_box, compare_meshes, last_pass_stats,  # noqa: F401
                   refine_by_size, simplex_degree)
from .osh_file import read_osh, write_osh  # noqa: F401
