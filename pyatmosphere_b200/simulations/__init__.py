"""Monte-Carlo driver and result containers (API of /root/reference/pyatmosphere/simulations/__init__.py:1-9 for
the records in scope)."""
from . import beam as _beam, measure as _measure, pdt as _pdt, result as _result, si as _si, simulation as _simulation, wind as _wind

Measure = _measure.Measure
Result = _result.Result
Simulation = _simulation.Simulation
BeamResult = _beam.BeamResult
PDTResult, TrackedPDTResult = _pdt.PDTResult, _pdt.TrackedPDTResult
SIResult = _si.SIResult
WindResult, TimeCoherenceResult, TimeBWcorrSimulation = _wind.WindResult, _wind.TimeCoherenceResult, _wind.TimeBWcorrSimulation

__all__ = ["Measure", "Result", "Simulation", "BeamResult", "PDTResult", "TrackedPDTResult", "SIResult", "WindResult",
           "TimeCoherenceResult", "TimeBWcorrSimulation"]
