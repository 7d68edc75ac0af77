from .result import Result
from .simulation import Simulation
from .measure import Measure

from .beam import BeamResult
from .pdt import PDTResult, TrackedPDTResult
