"""go1_gym_learn.ppo_cse — Runner of the reference (go1_gym_learn/ppo_cse/__init__.py:44-308) on the B200 kernels.

`Runner(env, device).learn(num_learning_iterations, init_at_random_ep_len, eval_freq)` keeps the reference's
loop structure (24-step rollout -> compute_returns -> update -> logging/checkpoints with the same file names),
so scripts/train.py drops in unchanged.  Multi-GPU: launch one process per GPU with torchrun; each rank owns
`env.num_envs` envs, and PPO gradients / advantage statistics / KL are all-reduced over NCCL
(`Runner(..., process_group=...)` or automatically when torch.distributed is initialised)."""
import copy
import os
import time
from collections import deque

import torch
from ml_logger import logger
from params_proto import PrefixProto

from .actor_critic import ActorCritic
from .rollout_storage import RolloutStorage


def class_to_dict(obj) -> dict:
    if not hasattr(obj, "__dict__"):
        return obj
    result = {}
    for key in dir(obj):
        if key.startswith("_") or key == "terrain":
            continue
        val = getattr(obj, key)
        result[key] = [class_to_dict(i) for i in val] if isinstance(val, list) else class_to_dict(val)
    return result


class DataCaches:
    def __init__(self, curriculum_bins):
        from go1_gym_learn.ppo.metrics_caches import SlotCache, DistCache
        self.slot_cache = SlotCache(curriculum_bins)
        self.dist_cache = DistCache()


caches = DataCaches(1)


class RunnerArgs(PrefixProto, cli=False):
    # runner
    algorithm_class_name = 'RMA'
    num_steps_per_env = 24  # per iteration
    max_iterations = 1500  # number of policy updates

    # logging
    save_interval = 400  # check for potential saves every this many iterations
    save_video_interval = 100
    log_freq = 10

    # load and resume
    resume = False
    load_run = -1  # -1 = last run
    checkpoint = -1  # -1 = last saved model
    resume_path = None  # updated from load_run and chkpt
    resume_curriculum = True


class Runner:

    def __init__(self, env, device='cpu', process_group=None):
        from .ppo import PPO

        self.device = device
        self.env = env
        actor_critic = ActorCritic(self.env.num_obs, self.env.num_privileged_obs, self.env.num_obs_history,
                                   self.env.num_actions).to(self.device)

        if RunnerArgs.resume:
            # the reference downloads from a fixed MIT server (ppo_cse/__init__.py:76-82); here resume_path is a
            # local directory written by a previous run (same file names).
            weights = torch.load(os.path.join(RunnerArgs.resume_path, "checkpoints", "ac_weights_last.pt"), map_location=self.device)
            actor_critic.load_state_dict(state_dict=weights)
            if hasattr(self.env, "curricula") and RunnerArgs.resume_curriculum:
                import pickle
                p = os.path.join(RunnerArgs.resume_path, "curriculum", "distribution.pkl")
                if os.path.exists(p):
                    rows = []
                    with open(p, "rb") as f:
                        while True:
                            try:
                                rows.append(pickle.load(f))
                            except EOFError:
                                break
                    distribution_last = rows[-1]["distribution"]
                    for gait_id, gait_name in enumerate(self.env.category_names):
                        self.env.curricula[gait_id].weights = distribution_last[f"weights_{gait_name}"]

        self.alg = PPO(actor_critic, device=self.device)
        self.num_steps_per_env = RunnerArgs.num_steps_per_env
        self.alg.init_storage(self.env.num_train_envs, self.num_steps_per_env, [self.env.num_obs],
                              [self.env.num_privileged_obs], [self.env.num_obs_history], [self.env.num_actions])
        if process_group is None and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            process_group = torch.distributed.group.WORLD
        self.process_group = process_group
        self.alg.process_group = process_group
        self.alg.storage.process_group = process_group
        if process_group is not None:       # identical initial weights on every rank
            torch.distributed.broadcast(actor_critic.flat_params, src=0, group=process_group)
            actor_critic.sample_seed = torch.distributed.get_rank(process_group)

        self.tot_timesteps = 0
        self.tot_time = 0
        self.current_learning_iteration = 0
        self.last_recording_it = 0
        self.collection_time = self.learn_time = 0.0

        self.env.reset()

    def rollout(self, obs, privileged_obs, obs_history, eval_expert=False):
        """The 24-step collection phase of learn() (ppo_cse/__init__.py:138-187)."""
        num_train_envs = self.env.num_train_envs
        infos = {}
        with torch.inference_mode():
            for i in range(self.num_steps_per_env):
                actions_train = self.alg.act(obs[:num_train_envs], privileged_obs[:num_train_envs], obs_history[:num_train_envs])
                if self.env.num_eval_envs > 0:
                    if eval_expert:
                        actions_eval = self.alg.actor_critic.act_teacher(obs_history[num_train_envs:], privileged_obs[num_train_envs:])
                    else:
                        actions_eval = self.alg.actor_critic.act_student(obs_history[num_train_envs:])
                    actions = torch.cat((actions_train, actions_eval), dim=0)
                else:
                    actions = actions_train
                obs_dict, rewards, dones, infos = self.env.step(actions)
                obs, privileged_obs, obs_history = obs_dict["obs"], obs_dict["privileged_obs"], obs_dict["obs_history"]
                self.alg.process_env_step(rewards[:num_train_envs], dones[:num_train_envs], infos)
                for key in ('train/episode', 'eval/episode'):
                    if key in infos:
                        if hasattr(logger, "store_metrics_lazy"):      # expanded once per log_metrics_summary, not per env step
                            logger.store_metrics_lazy(key, infos[key])
                        else:
                            with logger.Prefix(metrics=key):
                                logger.store_metrics(**infos[key])
        return obs, privileged_obs, obs_history, infos

    def learn(self, num_learning_iterations, init_at_random_ep_len=False, eval_freq=100, curriculum_dump_freq=500, eval_expert=False):
        from ml_logger import logger
        assert logger.prefix, "you will overwrite the entire instrument server"
        logger.start('start', 'epoch', 'episode', 'run', 'step')

        if init_at_random_ep_len:
            self.env.episode_length_buf = torch.randint_like(self.env.episode_length_buf, high=int(self.env.max_episode_length))

        num_train_envs = self.env.num_train_envs
        obs_dict = self.env.get_observations()
        obs, privileged_obs, obs_history = obs_dict["obs"], obs_dict["privileged_obs"], obs_dict["obs_history"]
        self.alg.actor_critic.train()

        rank0 = self.process_group is None or torch.distributed.get_rank(self.process_group) == 0
        tot_iter = self.current_learning_iteration + num_learning_iterations
        it = self.current_learning_iteration
        for it in range(self.current_learning_iteration, tot_iter):
            start = time.time()
            obs, privileged_obs, obs_history, infos = self.rollout(obs, privileged_obs, obs_history, eval_expert)
            distribution = infos.get('curriculum/distribution')
            stop = time.time()
            self.collection_time = stop - start
            start = stop
            with torch.inference_mode():
                self.alg.compute_returns(obs_history[:num_train_envs], privileged_obs[:num_train_envs])
            if it % curriculum_dump_freq == 0 and rank0:
                logger.save_pkl({"iteration": it, **caches.slot_cache.get_summary(), **caches.dist_cache.get_summary()},
                                path=f"curriculum/info.pkl", append=True)
                if distribution is not None:
                    logger.save_pkl({"iteration": it, "distribution": distribution}, path=f"curriculum/distribution.pkl", append=True)

            (mean_value_loss, mean_surrogate_loss, mean_adaptation_module_loss, mean_decoder_loss, mean_decoder_loss_student,
             mean_adaptation_module_test_loss, mean_decoder_test_loss, mean_decoder_test_loss_student) = self.alg.update()
            self.learn_time = time.time() - start

            logger.store_metrics(
                time_elapsed=logger.since('start'), time_iter=logger.split('epoch'),
                adaptation_loss=mean_adaptation_module_loss, mean_value_loss=mean_value_loss,
                mean_surrogate_loss=mean_surrogate_loss, mean_decoder_loss=mean_decoder_loss,
                mean_decoder_loss_student=mean_decoder_loss_student, mean_decoder_test_loss=mean_decoder_test_loss,
                mean_decoder_test_loss_student=mean_decoder_test_loss_student,
                mean_adaptation_module_test_loss=mean_adaptation_module_test_loss)

            if RunnerArgs.save_video_interval:
                self.log_video(it)

            world = 1 if self.process_group is None else torch.distributed.get_world_size(self.process_group)
            self.tot_timesteps += self.num_steps_per_env * self.env.num_envs * world
            if logger.every(RunnerArgs.log_freq, "iteration", start_on=1) and rank0:
                logger.log_metrics_summary(key_values={"timesteps": self.tot_timesteps, "iterations": it})
                logger.job_running()

            if it % RunnerArgs.save_interval == 0 and rank0:
                self.save(it)
            self.current_learning_iteration += num_learning_iterations
        if rank0:
            self.save(it)

    def save(self, it):
        """Same artefacts as the reference (ppo_cse/__init__.py:231-274): ac_weights_{it:06d}.pt (+ _last), and
        TorchScript exports of the adaptation module and actor body for go1_gym_deploy."""
        with logger.Sync():
            logger.torch_save(self.alg.actor_critic.state_dict(), f"checkpoints/ac_weights_{it:06d}.pt")
            logger.duplicate(f"checkpoints/ac_weights_{it:06d}.pt", f"checkpoints/ac_weights_last.pt")
            path = './tmp/legged_data'
            os.makedirs(path, exist_ok=True)
            adaptation_module_path = f'{path}/adaptation_module_latest.jit'
            adaptation_module = copy.deepcopy(self.alg.actor_critic.adaptation_module).to('cpu')
            torch.jit.script(adaptation_module).save(adaptation_module_path)
            body_path = f'{path}/body_latest.jit'
            body_model = copy.deepcopy(self.alg.actor_critic.actor_body).to('cpu')
            torch.jit.script(body_model).save(body_path)
            logger.upload_file(file_path=adaptation_module_path, target_path=f"checkpoints/", once=False)
            logger.upload_file(file_path=body_path, target_path=f"checkpoints/", once=False)

    def log_video(self, it):
        if it - self.last_recording_it >= RunnerArgs.save_video_interval:
            self.env.start_recording()
            if self.env.num_eval_envs > 0:
                self.env.start_recording_eval()
            self.last_recording_it = it
        frames = self.env.get_complete_frames()
        if len(frames) > 0:
            self.env.pause_recording()
            logger.save_video(frames, f"videos/{it:05d}.mp4", fps=1 / self.env.dt)
        if self.env.num_eval_envs > 0:
            frames = self.env.get_complete_frames_eval()
            if len(frames) > 0:
                self.env.pause_recording_eval()
                logger.save_video(frames, f"videos/{it:05d}_eval.mp4", fps=1 / self.env.dt)

    def get_inference_policy(self, device=None):
        self.alg.actor_critic.eval()
        if device is not None:
            self.alg.actor_critic.to(device)
        return self.alg.actor_critic.act_inference

    def get_expert_policy(self, device=None):
        self.alg.actor_critic.eval()
        if device is not None:
            self.alg.actor_critic.to(device)
        return self.alg.actor_critic.act_expert
