"""Vehicle parameter sets the driver selects by name (planning.py:297-298 reads
``commonroad_dc...VehicleParameterMapping[VehicleType.VW_VANAGON.name].value``; package
``commonroad-vehicle-models`` 3.0.2 -- third-party, not in the reference tree).

Only ``l, w, a, b, T_f, T_r, longitudinal.{v_max, a_max}, steering.{max, v_max, kappa_dot_max, kappa_dot_dot_max}``
are read (vehicle.py:17-46), and only ``l, w, longitudinal.v_max, longitudinal.a_max`` reach the active path.
In-tree pins: footprint 4.569 x 1.844 m (SMP/maneuver_automaton/maneuver_automaton.py:47-48), steering limit
1.023 (automaton file names).  The remaining numbers are the published vehicle-model values restated from memory
[flagged in SURVEY 8(c)(ii)]; pass your own namespace to ``frenet_optimal_planning`` to override them.
"""
import enum
import types


def _params(l, w, a, b, T_f, T_r, v_max, a_max, steer_max, steer_v_max, kappa_dot_max, kappa_dot_dot_max):
    return types.SimpleNamespace(
        l=l, w=w, a=a, b=b, T_f=T_f, T_r=T_r,
        longitudinal=types.SimpleNamespace(v_max=v_max, a_max=a_max),
        steering=types.SimpleNamespace(max=steer_max, min=-steer_max, v_max=steer_v_max, v_min=-steer_v_max,
                                       kappa_dot_max=kappa_dot_max, kappa_dot_dot_max=kappa_dot_dot_max))


class VehicleType(enum.Enum):
    FORD_ESCORT = 1
    BMW_320i = 2
    VW_VANAGON = 3


class VehicleParameterMapping(enum.Enum):
    FORD_ESCORT = _params(4.298, 1.674, 1.0203, 1.5547, 1.431, 1.4494, 45.8, 11.5, 0.910, 0.4, 0.4, 20.0)
    BMW_320i = _params(4.508, 1.610, 1.1562, 1.4227, 1.4427, 1.4173, 50.8, 11.5, 1.066, 0.4, 0.4, 20.0)
    VW_VANAGON = _params(4.569, 1.844, 1.1508, 1.3211, 1.5745, 1.5311, 41.7, 11.5, 1.023, 0.4, 0.4, 20.0)
