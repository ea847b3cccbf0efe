"""BASELINE config 1 / the reference's CI known answer (.github/workflows/train-example.yml:26-27):

    python src/main.py --model=dqn --random-topology=1 --mini-batch-size=32 --episode-steps=1
        --eval-episode-steps=1 --lr=0.001 --tau=0.01 --netmon --netmon-encoder-dim=4 --hidden-dim=4
        --netmon-dim=2 --netmon-iterations=1 --sequence-length=1 --step-before-train=1_000
        --capacity=10_000 --eval-episodes=100 --total-steps=5_000 --env-type=simple --epsilon=0.1
        --epsilon-decay=1.0 --seed=0            ->  "reward_mean": 1.0

This driver is the rollout + learner loop of src/main.py:667-1026 and the evaluation of
src/eval.py:33-168 with the classes of graph_marl_b200 dropped in (rollout on the CUDA kernels in
compat mode, learner through the modules' autograd path).  Prints a JSON line with reward_mean.
"""
import argparse
import copy
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from graph_marl_b200.env.environment import reset_and_get_sizes  # noqa: E402
from graph_marl_b200.env.simple_environment import SimpleEnvironment  # noqa: E402
from graph_marl_b200.env.wrapper import NetMonWrapper  # noqa: E402
from graph_marl_b200.model import DQN, NetMon  # noqa: E402
from graph_marl_b200.policy import EpsilonGreedy  # noqa: E402
from graph_marl_b200.replaybuffer import ReplayBuffer  # noqa: E402
from graph_marl_b200.util import interpolate_model, set_seed  # noqa: E402


def evaluate(env, policy, episodes, steps_per_episode):
    """eval.py:33-168 reduced to the metric means."""
    policy.eval()
    rewards = []
    for _ in range(episodes):
        obs, adj = env.reset()
        policy.reset(1)
        for _ in range(steps_per_episode):
            actions = policy(obs, adj)
            obs, adj, reward, done, info = env.step(actions)
            policy.reset(done)
            rewards.append(np.mean(reward))
    return {"reward_mean": float(np.mean(rewards))}


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--total-steps", type=int, default=5000)
    ap.add_argument("--step-before-train", type=int, default=1000)
    ap.add_argument("--eval-episodes", type=int, default=100)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--device", default="cuda")
    a = ap.parse_args(argv)
    args = SimpleNamespace(epsilon=0.1, epsilon_decay=1.0, epsilon_update_freq=100, step_before_train=a.step_before_train,
                           mini_batch_size=32, sequence_length=1, gamma=0.98, lr=1e-3, tau=0.01, capacity=10000,
                           episode_steps=1, step_between_train=1)
    set_seed(a.seed)
    dev = torch.device(a.device)
    env = SimpleEnvironment(env_var=1, random_topology=1, device=dev)
    n_agents, agent_obs_size, n_nodes, node_obs_size = reset_and_get_sizes(env)
    netmon = NetMon(node_obs_size, 2, [4], 1, F.leaky_relu, rnn_type="lstm", rnn_carryover=True, agg_type="sum",
                    output_neighbor_hidden=True, output_global_hidden=False).to(dev)
    node_state_size = netmon.get_state_size()
    env = NetMonWrapper(env, netmon, 1)
    n_agents, obs_size, _, _ = reset_and_get_sizes(env)  # main.py:478 sizes the DQN from a live reset
    model = DQN(obs_size, [4], env.action_space.n, F.leaky_relu).to(dev)
    model_tar = copy.deepcopy(model).to(dev)
    policy = EpsilonGreedy(env, model, env.action_space.n, args)
    parameters = list(model.parameters()) + list(netmon.parameters())
    optimizer = torch.optim.AdamW(parameters, lr=args.lr)
    buff = ReplayBuffer(a.seed, args.capacity, n_agents, obs_size, 0, n_nodes, node_obs_size, node_state_size, 0, device=dev)

    episode_step, episode_done = None, False
    for step in range(1, a.total_steps + 1):
        model.eval(), netmon.eval()
        if episode_step is None or episode_done:
            episode_step = 0
            obs, adj = env.reset()
        buffer_node_state = env.last_netmon_state.cpu().detach().numpy() if env.last_netmon_state is not None else 0
        netmon_info = env.get_netmon_info()
        joint_actions = policy(obs, adj)
        next_obs, next_adj, reward, done, info = env.step(joint_actions)
        next_netmon_info = env.get_netmon_info()
        episode_step += 1
        episode_done = episode_step >= args.episode_steps
        buff.add(obs, joint_actions, reward, next_obs, adj, next_adj, done, episode_done, 0, buffer_node_state, None,
                 *netmon_info, *next_netmon_info)
        obs, adj = next_obs, next_adj
        if step < args.step_before_train or buff.count < args.mini_batch_size or step % args.step_between_train != 0:
            continue
        # ---- learner (main.py:830-1026), sequence_length 1 -----------------------------------
        model.train(), netmon.train()
        loss_q = torch.zeros(1, device=dev)
        for t, batch in enumerate(buff.get_batch(args.mini_batch_size, device=dev, sequence_length=args.sequence_length)):
            netmon.state = batch.node_state
            network_obs = netmon(batch.node_obs, batch.node_adj, batch.node_agent_matrix)
            net_obs_dim = network_obs.shape[-1]
            b_obs = torch.cat((batch.obs[:, :, :-net_obs_dim], network_obs), dim=-1)  # main.py:879-880 (in-place there)
            q_values = model(b_obs, batch.adj)
            with torch.no_grad():
                next_network_obs = netmon(batch.next_node_obs, batch.next_node_adj, batch.next_node_agent_matrix)
                b_next = torch.cat((batch.next_obs[:, :, :-next_network_obs.shape[-1]], next_network_obs), dim=-1)
                next_q_max = model_tar(b_next, batch.next_adj).max(dim=2)[0]
            target = batch.reward + (~batch.done) * args.gamma * next_q_max
            q_target = torch.scatter(q_values.detach(), -1, batch.action.unsqueeze(-1), target.unsqueeze(-1))
            loss_q = loss_q + torch.mean((q_values - q_target).pow(2)) / args.sequence_length
        optimizer.zero_grad()
        loss_q.backward()
        torch.nn.utils.clip_grad_value_(parameters, 0.5)
        torch.nn.utils.clip_grad_norm_(parameters, 1.0)
        optimizer.step()
        interpolate_model(model, model_tar, args.tau, model_tar)

    model.eval(), netmon.eval()
    metrics = evaluate(env, policy, a.eval_episodes, 1)
    print(json.dumps(metrics, sort_keys=True))
    return metrics


if __name__ == "__main__":
    main()
