"""yacs-free stand-in for the reference's configuration objects (softmac/config/*.py).

``CfgNode`` keeps the attribute names of cfg.SIMULATOR / cfg.PRIMITIVES / cfg.RIGID
(softmac/config/default_config.py:14-60) so code written against the reference configs reads the same."""


class CfgNode(dict):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        for k, v in list(self.items()):
            if isinstance(v, dict) and not isinstance(v, CfgNode):
                self[k] = CfgNode(v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        import copy
        return copy.deepcopy(self)

    def defrost(self):
        pass

    def freeze(self):
        pass


CN = CfgNode


def simulator_defaults():
    """cfg.SIMULATOR defaults, softmac/config/default_config.py:14-29."""
    return CfgNode(dim=3, quality=1, yield_stress=50., dtype="float64", max_steps=1024, n_particles=9000, E=5e3, nu=0.2,
                   ground_friction=1.5, gravity=(0, 0, 0), ptype=0, material_model=1, dt=1e-4, n_controllers=0,
                   collision_type=2)


def get_cfg_defaults():
    return CfgNode(control_mode="rigid", rigid_velocity_control=False, env_dt=2e-3, SIMULATOR=simulator_defaults(),
                   PRIMITIVES=[], SHAPES=[], RIGID=CfgNode(gravity=(0., 0., 0.), init_state=(), enable_floor=True),
                   ENV=CfgNode(loss_type="", loss=CfgNode(soft_contact=False, weight=(10., 10., 1.), target_path=""),
                               n_observed_particles=200), VARIANTS=[])
