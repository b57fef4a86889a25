"""DTCSimulator as a qiskit BackendV2 (fast.py:181-191): qiskit is not installed here, so stub modules with
the constructor / method surface the factory touches are injected into sys.modules to prove the branch constructs and
offers what `generate_preset_pass_manager(backend=backend, ...)` and the scripts read."""
import sys
import types

import pytest


class _Gate:
    def __init__(self, *params):
        self.params = list(params)
        self.name = type(self).__name__.replace("Gate", "").lower()


def _stub_qiskit():
    mods = {}
    qk = types.ModuleType("qiskit")
    circuit = types.ModuleType("qiskit.circuit")
    library = types.ModuleType("qiskit.circuit.library")
    providers = types.ModuleType("qiskit.providers")
    transpiler = types.ModuleType("qiskit.transpiler")

    class Parameter:
        def __init__(self, name):
            self.name = name

    class Measure(_Gate):
        pass

    for nm in ("CXGate", "IGate", "RZGate", "SXGate", "U1Gate", "U2Gate", "U3Gate"):
        setattr(library, nm, type(nm, (_Gate,), {}))
    library.IGate.__init__ = lambda self: (_Gate.__init__(self), setattr(self, "name", "id"))[0]

    class Options(dict):
        def __init__(self, **kw):
            super().__init__(**kw)

    class BackendV2:
        def __init__(self, provider=None, name=None, description=None, online_date=None, backend_version=None, **fields):
            self._provider, self.name, self.description, self.backend_version = provider, name, description, backend_version
            self._options = self._default_options()

        @property
        def options(self):
            return self._options

        @property
        def operation_names(self):
            return list(self.target.operation_names)

    class Target:
        def __init__(self, num_qubits=None, description=None):
            self.num_qubits, self.description, self._ops = num_qubits, description, {}

        def add_instruction(self, instruction, properties=None, name=None):
            self._ops[name or instruction.name] = properties

        @property
        def operation_names(self):
            return list(self._ops)

    circuit.Parameter, circuit.Measure = Parameter, Measure
    providers.BackendV2, providers.Options = BackendV2, Options
    transpiler.Target = Target
    qk.circuit, qk.providers, qk.transpiler = circuit, providers, transpiler
    circuit.library = library
    mods.update({"qiskit": qk, "qiskit.circuit": circuit, "qiskit.circuit.library": library,
                 "qiskit.providers": providers, "qiskit.transpiler": transpiler})
    return mods, BackendV2


def test_backend_v2_branch_constructs(monkeypatch):
    from dtcsim import backend
    mods, BackendV2 = _stub_qiskit()
    for k, v in mods.items():
        monkeypatch.setitem(sys.modules, k, v)
    cls = backend.make_backendv2_class(backend.DTCSimulatorBase)
    sim = cls(noise_model=None, device="GPU", cuStateVec_enable=True)        # the reference's constructor call, fast.py:156
    assert isinstance(sim, BackendV2) and isinstance(sim, backend.DTCSimulatorBase)
    assert sim.name == "aer_simulator"                                       # file names embed it, fast.py:191,196
    t = sim.target
    assert sorted(t.operation_names) == sorted(backend.TARGET_OPERATIONS)
    assert t.num_qubits >= 31 and all(v is None for v in t._ops.values())    # all-to-all, ideal (snake layout reaches 30)
    assert sim.max_circuits is None and sim.options["shots"] == 1024
    assert cls.run.__qualname__.startswith("make_backendv2_class")           # run() is ours, not the abstract one
    sim.set_options(shots=77, foo=1)
    assert sim.default_shots == 77 and sim.sim_options["foo"] == 1
    # the lowering mirror accepts it as backend= as well (physical width from the backend)
    import dtcsim
    pm = dtcsim.generate_preset_pass_manager(optimization_level=0, backend=sim, initial_layout=dtcsim.SNAKE_LAYOUT[:5],
                                             routing_method=None)
    c = dtcsim.QuantumCircuit(5, 1)
    c.h(0)
    c.measure(0, 0)
    assert pm.run(c).num_qubits >= 31


def test_without_qiskit_plain_class_is_exported():
    from dtcsim import backend
    try:
        import qiskit  # noqa: F401
    except ImportError:
        assert backend.DTCSimulator is backend.DTCSimulatorBase
        with pytest.raises(ImportError):
            backend.make_backendv2_class()
