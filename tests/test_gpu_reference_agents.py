"""The reference's OWN agents, unmodified (imported from baseline/_ref, never copied), driving the drop-in env.

north_star: "the finenvs/environments reset()/step(actions) interface stays a drop-in ... so finenvs/agents (PPO, ES)
run unchanged".  These tests run the two canonical loops of the reference with nothing replaced but the env class:

  * examples/time_series/PPO_LSTM_training_SPY.py:22-30 — PPOAgentLSTM (finenvs/agents/PPO/PPO_agent.py:246-285):
    step -> env.step -> store -> train(); needs obs (N, W, 5) that `.float()` accepts, f32-addable rewards, integer
    0/1 dones with `dones[-1].item()`, a FRESH obs tensor per step (the reference Buffer torch.cat's them), and the
    "last env is the evaluation env" convention;
  * examples/isaac_gym/ES_MLP_Isaac_Gym.py:30-38 — EvoAgent + ParallelMLP (finenvs/agents/ES/evo_agent.py,
    finenvs/agents/networks/parallel_mlp.py): needs (N, num_observations) fp32 observations (flat_obs=True),
    env_args["num_eval_envs"], dones.nonzero(), env.reset_all().

Skipped where baseline/_ref does not exist (`python __graft_entry__.py` creates it wherever /root/reference is;
the tree is git-ignored but travels to the GPU box).
"""
import os

import pytest
import torch

from baseline import reference as ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="baseline/_ref (the reference tree) is absent")]


@pytest.fixture()
def reference_tree(tmp_path, monkeypatch):
    ref.import_reference()
    monkeypatch.setenv("FINENVS_DATA_DIR", os.path.join(ref.REF_DST, "finenvs", "data"))   # the reference's own CSVs
    monkeypatch.chdir(tmp_path)   # the agents create ./trials
    return ref


def test_reference_ppo_lstm_agent_trains_on_the_dropin(reference_tree):
    from finenvs.agents.PPO.PPO_agent import PPOAgentLSTM   # the reference's class, from baseline/_ref

    from finenvs_b200.environments import TimeSeriesEnv

    assert __import__("finenvs").__file__.startswith(ref.REF_DST)
    torch.manual_seed(602585557)                                      # PPO_LSTM_training_SPY.py:34
    env = TimeSeriesEnv(instrument_name="SPY", dataset_key="dummy", num_intervals=4)   # :10 ("train" is not distributed)
    env_args = env.get_env_args()
    num_envs = env_args["num_envs"]
    batch_size = num_envs * 64
    agent = PPOAgentLSTM(env_args=env_args, num_epochs=4, num_mini_batches=4, hidden_dim=1024, model_save_interval=0,
                         write_to_csv=False)
    states = env.reset()
    assert states.shape == (num_envs, 4, 5)
    trained, last_samples = 0, None
    for it in range(150):                                             # :22-30, two train() calls at 64 steps per batch
        (actions, log_probs, values) = agent.step(states)
        held = states.clone()
        (next_states, rewards, dones, _) = env.step(actions)
        assert torch.equal(states, held)                              # step() returns a fresh tensor: the observation in hand
        agent.store(states, actions, rewards, dones, log_probs, values)   # (and those the reference Buffer cat'ed) stay intact
        states = next_states
        if agent.get_buffer_size() >= batch_size:
            before = [p.detach().clone() for p in agent.actor.parameters()]
            last_samples = agent.train(states)
            trained += 1
            assert any(not torch.equal(a, b) for a, b in zip(before, agent.actor.parameters())), "train() changed nothing"
            assert agent.get_buffer_size() == 0
    assert trained == 2
    assert all(torch.isfinite(p).all() for p in agent.actor.parameters())
    assert all(torch.isfinite(p).all() for p in agent.critic.parameters())
    assert torch.isfinite(agent.current_returns).all()
    assert agent.num_samples == 2 * batch_size and last_samples in (0, agent.num_samples)


def test_reference_es_agent_trains_on_the_dropin(reference_tree):
    from finenvs.agents.ES.evo_agent import EvoAgent   # the reference's class, from baseline/_ref
    from finenvs.agents.networks.parallel_mlp import ParallelMLP

    from finenvs_b200.environments import TimeSeriesEnv

    torch.manual_seed(0)                                              # ES_MLP_Isaac_Gym.py:9
    num_envs, num_eval_envs, episodes_per_batch = 1036, 12, 2500
    env = TimeSeriesEnv("SPY", "dummy", num_intervals=4, num_envs=num_envs, flat_obs=True, num_eval_envs=num_eval_envs,
                        random_reset="all", seed=3)
    env_args = env.get_env_args()
    assert env_args["num_observations"] == 20 and env_args["num_eval_envs"] == num_eval_envs
    agent = EvoAgent(env_args, hidden_dims=(256, 256), learning_rate=0.01, noise_std_dev=0.02, l2_coefficient=0.005,
                     write_to_csv=False)                              # :22-29
    assert isinstance(agent.network, ParallelMLP)
    theta0 = [w.clone() for w in agent.network.weight_layers]
    states = env.reset()
    assert states.shape == (num_envs, 20) and states.dtype == torch.float32
    generations = 0
    for it in range(4000):                                            # :30-38
        actions = agent.step(states)
        (next_states, rewards, dones, _) = env.step(actions)
        (num_done, _) = agent.store(rewards, dones)
        states = next_states
        if num_done >= episodes_per_batch:
            agent.train()
            states = env.reset_all()
            generations += 1
            if generations == 2:
                break
    assert generations == 2
    assert any(not torch.equal(a, b) for a, b in zip(theta0, agent.network.weight_layers))
    assert all(torch.isfinite(w).all() for w in agent.network.weight_layers)
