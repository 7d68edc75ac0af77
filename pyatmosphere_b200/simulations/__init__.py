from .result import Result
from .simulation import Simulation
from .measure import Measure

from .beam import BeamResult
from .si import SIResult
from .pdt import PDTResult, TrackedPDTResult
from .wind import WindResult, TimeCoherenceResult, TimeBWcorrSimulation
