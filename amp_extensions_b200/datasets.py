"""Offline (s, a, s') dataset with the reference's normalisation statistics.

Mirror of reference milo/milo/datasets.py:8-49 (AmpDataset): same constructor, same
`get_transformations` definition (mean and mean-absolute-deviation + 1e-8), same item layout, so it is
accepted wherever the reference class is (DataLoader, DynamicsEnsemble ctor).
"""
import torch
from torch.utils.data import Dataset


class AmpDataset(Dataset):
    def __init__(self, states, actions, next_states, device=torch.device("cpu")):
        self.device = device
        self.states = states
        self.actions = actions
        self.next_states = next_states

    def get_transformations(self, device=None):
        """(state_mean, state_scale, action_mean, action_scale, diff_mean, diff_scale); datasets.py:23-43."""
        diff = self.next_states - self.states
        state_mean = self.states.mean(dim=0).float()
        action_mean = self.actions.mean(dim=0).float()
        diff_mean = diff.mean(dim=0).float()
        state_scale = (self.states - state_mean).abs().mean(dim=0).float() + 1e-8
        action_scale = (self.actions - action_mean).abs().mean(dim=0).float() + 1e-8
        diff_scale = (diff - diff_mean).abs().mean(dim=0).float() + 1e-8
        dev = self.device if device is None else device
        return tuple(t.to(dev) for t in (state_mean, state_scale, action_mean, action_scale, diff_mean, diff_scale))

    def __len__(self):
        return self.states.size(0)

    def __getitem__(self, idx):
        return self.states[idx].float(), self.actions[idx].float(), self.next_states[idx].float()
