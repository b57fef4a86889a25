"""dtcsim -- B200-native simulator for the reference's noisy kicked-Ising DTC Floquet circuits.

Drop-in for the ``AerSimulator(noise_model=...).run(circ, shots).result().get_counts()`` path of
the reference's ``autocorr-*-qiskit-fast*.py`` scripts (fast.py:156,211-212).  Host side in Python
(mirroring the reference's Python boundary), compute in hand-written sm_100a CUDA reached through
the C ABI in ``include/dtcsim.h``; PyTorch only allocates device buffers and provides streams.
"""
from .ir import QuantumCircuit, Op, from_qasm2, from_qiskit, as_circuit          # noqa: F401
from .lowering import generate_preset_pass_manager, lower_level0, SNAKE_LAYOUT   # noqa: F401
from .noise import (NoiseModel, ReadoutError, ChannelError, depolarizing_error, pauli_error, kraus_error,   # noqa: F401
                    amplitude_damping_error, phase_damping_error, thermal_relaxation_error, as_noise_model)
from .plan import compile_circuit, Program                                       # noqa: F401
from .sweeps import (autocorr_circuit, energy_circuit, expz_circuit, feedback_g, floquet_period, optimize_g,   # noqa: F401
                     run_adaptive, run_energy_sweep, run_expz_sweep, run_shots_sweep, run_sweep, xy_cycle_schedule)
from .estimator import BackendEstimatorV2, BackendSamplerV2, SamplerV2, dtc_hamiltonian   # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # the backend needs torch + the CUDA library; import lazily so host-only tooling stays light
    if name in ("DTCSimulator", "AerSimulator", "Job", "Result", "compute_z_expectation"):
        from . import backend
        return getattr(backend, name)
    raise AttributeError(name)
