"""GAIL / AMP discriminator cost evaluated on B200 (reference milo/milo/gail_cost.py:45-283).

Training the discriminator (update_disc: least-squares / log-likelihood losses, gradient penalty, optimiser) stays
with the reference class on the host — it is outside the learned-dynamics step.  What
the rollout consumes, `get_costs(ss)` and `get_bonus_costs(states, actions, ensemble, next_states)`
(mjrl/mjrl/algos/batch_reinforce.py:128, 146-166), runs through libsimstep: the discriminator's hidden layers in
the grouped tcgen05 GEMM, its last layer as a linear head, the AMP cost transform and the pessimism bonus in the
combine kernel.

    ref = milo.gail_cost.GAILCost(expert_ss, ...)        # the reference object: owns disc, optimiser, update_disc
    cost = amp_extensions_b200.GAILCost(ref)              # same attributes (delegated), device get_costs
"""
import torch
import torch.nn as nn

from . import engine as _engine


class GAILCost:
    def __init__(self, reference_cost, precision=None, device=None):
        """reference_cost: an object with the reference GAILCost's attributes `disc` (Discriminator with `.net`),
        `lambda_b`, `input_type`, `disc_loss_type` (gail_cost.py:45-93)."""
        object.__setattr__(self, "_ref", reference_cost)
        object.__setattr__(self, "_precision", precision)
        object.__setattr__(self, "_device", device)
        object.__setattr__(self, "_eng", None)
        object.__setattr__(self, "_stamp", None)

    # everything this wrapper does not define is the reference object's (update_disc, disc_opt, expert_data, ...)
    def __getattr__(self, name):
        return getattr(object.__getattribute__(self, "_ref"), name)

    def __setattr__(self, name, value):
        setattr(object.__getattribute__(self, "_ref"), name, value)

    def _linears(self):
        net = self._ref.disc.net
        return [net] if isinstance(net, nn.Linear) else [m for m in net if isinstance(m, nn.Linear)]

    def engine(self):
        lin = self._linears()
        if lin[-1].out_features != 1:
            raise NotImplementedError("device GAIL costs support a scalar discriminator output (feature_dim = 1)")
        if self._eng is None:
            act = self._ref.disc.activation
            eng = _engine.Engine(state_dim=lin[0].in_features, action_dim=0, num_models=1,
                                 hidden_sizes=[l.out_features for l in lin[:-1]], dense_connect=False,
                                 activation="relu" if isinstance(act, nn.ReLU) else "tanh", transform=False,
                                 precision=self._precision, device=self._device)
            object.__setattr__(self, "_eng", eng)
            object.__setattr__(self, "_ones", torch.ones(1, device=eng.device))
        stamp = tuple((l.weight._version, l.bias._version, id(l.weight.data), id(l.bias.data)) for l in lin)
        if stamp != self._stamp:   # the optimiser updates the parameters in place: re-pack after update_disc
            self._eng.load_feature_net([l.weight.data for l in lin[:-1]], [l.bias.data for l in lin[:-1]],
                                       lin[-1].weight.data, lin[-1].bias.data, head_mode=_engine.Engine.HEAD_LINEAR)
            object.__setattr__(self, "_stamp", stamp)
        self._eng.set_cost_transform(_engine.Engine.COST_GAIL_LS if self._ref.disc_loss_type == "least_squares"
                                     else _engine.Engine.COST_GAIL_LL)
        return self._eng

    @torch.no_grad()
    def disc_outputs(self, ss):
        """self.disc(ss) (gail_cost.py:42-43), CPU tensor [n, 1]."""
        eng = self.engine()
        eng.set_cost_transform(_engine.Engine.COST_IDENTITY)
        return eng.rff_dot(ss, self._ones).cpu().unsqueeze(1)

    @torch.no_grad()
    def get_costs(self, ss):
        """gail_cost.py:248-253: the AMP least-squares or log-likelihood cost, CPU tensor [n, 1]."""
        return self.engine().rff_dot(ss, self._ones).cpu().unsqueeze(1)

    @torch.no_grad()
    def get_bonus_costs(self, states, actions, ensemble, next_states=None):
        """gail_cost.py:255-283: cost = (1 - lambda_b) c(ss) - lambda_b * discrepancy (no threshold clamp)."""
        t = self._ref.input_type
        states, actions = states.float(), actions.float()
        if t == "sa":
            x = torch.cat([states, actions], dim=1)
        elif t == "ss":
            assert next_states is not None
            x = torch.cat([states, next_states.float()], dim=1)
        elif t == "sas":
            x = torch.cat([states, actions, next_states.float()], dim=1)
        elif t == "s":
            x = states
        else:
            raise NotImplementedError("Input type not implemented")
        eng = self.engine()
        disc = ensemble.get_action_discrepancy(states, actions)
        lam = self._ref.lambda_b
        cost, ipm, bonus = eng.bonus_cost(x, disc, self._ones, lam, 1.0, 0.0, 0.0, False)
        cost, ipm, bonus = cost.cpu().unsqueeze(1), ipm.cpu().unsqueeze(1), bonus.cpu().unsqueeze(1)
        v_targ = ipm / (1 - lam) if lam != 1 else self.get_costs(x)
        return cost, {"bonus": bonus, "ipm": ipm, "v_targ": v_targ, "cost": cost}
