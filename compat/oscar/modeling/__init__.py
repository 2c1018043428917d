"""`oscar.modeling`: modeling_vlbert is the B200 drop-in; other modules resolve to the reference checkout."""
from oscar import _extend

__version__ = "0.1.0"
_extend(__path__, ("oscar", "modeling"))
