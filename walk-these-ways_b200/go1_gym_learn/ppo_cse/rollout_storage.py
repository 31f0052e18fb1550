"""RolloutStorage of ppo_cse (reference go1_gym_learn/ppo_cse/rollout_storage.py:5-178): [T, N, .] slabs,
GAE through the warp-scan kernel, minibatches through the row-gather kernel."""
import torch

from go1_b200 import capi


class RolloutStorage:
    class Transition:
        def __init__(self):
            self.observations = None
            self.privileged_observations = None
            self.observation_histories = None
            self.critic_observations = None
            self.actions = None
            self.rewards = None
            self.dones = None
            self.values = None
            self.actions_log_prob = None
            self.action_mean = None
            self.action_sigma = None
            self.env_bins = None

        def clear(self):
            self.__init__()

    def __init__(self, num_envs, num_transitions_per_env, obs_shape, privileged_obs_shape, obs_history_shape, actions_shape, device='cpu'):
        self.device = device
        self.obs_shape, self.privileged_obs_shape = obs_shape, privileged_obs_shape
        self.obs_history_shape, self.actions_shape = obs_history_shape, actions_shape
        T, N = num_transitions_per_env, num_envs
        z = lambda *s, **k: torch.zeros(T, N, *s, device=self.device, **k)
        self.observations = z(*obs_shape)
        self.privileged_observations = z(*privileged_obs_shape)
        self.observation_histories = z(*obs_history_shape)
        # Row pitch of the GATHERED minibatch histories: a multiple of 32 floats, so that every 128-byte row of a TMA box of the first-layer
        # products starts on a 128-byte line.  Measured (tools/epi_bench.py): the 24576 x 1280 x 2100 product runs in 180 us with a
        # 2112-float pitch against 259 us with the natural 2100 (each misaligned box row costs a fifth L2 sector).
        self.hist_pitch = (int(obs_history_shape[-1]) + 31) // 32 * 32
        self.rewards = z(1)
        self.actions = z(*actions_shape)
        self.dones = z(1).byte()
        self.actions_log_prob = z(1)
        self.values = z(1)
        self.returns = z(1)
        self.advantages = z(1)
        self.mu = z(*actions_shape)
        self.sigma = z(*actions_shape)
        self.env_bins = z(1)
        self.num_transitions_per_env, self.num_envs = T, N
        self.step = 0
        self._stats = torch.zeros(2, dtype=torch.float64, device=self.device)
        self.process_group = None          # set by the multi-GPU runner: advantage statistics are global

    def add_transitions(self, transition: Transition):
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        t = self.step
        self.observations[t].copy_(transition.observations)
        self.privileged_observations[t].copy_(transition.privileged_observations)
        self.observation_histories[t].copy_(transition.observation_histories)
        self.actions[t].copy_(transition.actions)
        self.rewards[t].copy_(transition.rewards.view(-1, 1))
        self.dones[t].copy_(transition.dones.view(-1, 1))
        self.values[t].copy_(transition.values)
        self.actions_log_prob[t].copy_(transition.actions_log_prob.view(-1, 1))
        self.mu[t].copy_(transition.action_mean)
        self.sigma[t].copy_(transition.action_sigma)
        self.env_bins[t].copy_(transition.env_bins.view(-1, 1))
        self.step += 1

    def snapshot_observations(self, obs, privileged_obs):
        """Copy the observations the policy acts on into slot `step` NOW.  The env writes its next observations into the same
        buffers during env.step, so a reference kept until process_env_step would store s_{t+1} next to a_t (the reference is
        safe only because its compute_observations allocates fresh tensors).  Returns the two storage views."""
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        t = self.step
        so, sp = self.observations[t], self.privileged_observations[t]
        if obs.is_cuda and obs.dtype == torch.float32 and obs.is_contiguous() and privileged_obs.dtype == torch.float32 and privileged_obs.is_contiguous():
            capi.check(capi.lib().go1_store_observations(capi.ptr(obs), capi.ptr(privileged_obs) if sp.shape[-1] else None, capi.ptr(so),
                                                         capi.ptr(sp) if sp.shape[-1] else None, self.num_envs, so.shape[-1], sp.shape[-1],
                                                         capi.stream_ptr()), "go1_store_observations")
        else:
            so.copy_(obs); sp.copy_(privileged_obs)
        return so, sp

    def add_transitions_fused(self, tr, time_outs, gamma):
        """add_transitions + the time-out bootstrap in ONE kernel (go1_store_transition)."""
        import ctypes as C
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        t = self.step
        u8 = lambda x: x if x.dtype == torch.uint8 else (x.view(torch.uint8) if x.dtype == torch.bool else x.to(torch.uint8))
        dones, touts = u8(tr.dones), (u8(time_outs) if time_outs is not None else None)
        stored = lambda x, slot: None if (x is None or x.data_ptr() == slot.data_ptr()) else x       # already snapshotted by act()
        ins = [stored(tr.observations, self.observations[t]), stored(tr.privileged_observations, self.privileged_observations[t]),
               tr.observation_histories, tr.actions, tr.rewards, tr.values, tr.actions_log_prob, tr.action_mean, tr.action_sigma_vec, tr.env_bins]
        outs = [self.observations[t], self.privileged_observations[t], self.observation_histories[t], self.actions[t], self.rewards[t], self.values[t],
                self.actions_log_prob[t], self.mu[t], self.sigma[t], self.env_bins[t]]
        for x in ins:
            assert x is None or (x.is_contiguous() and x.dtype == torch.float32 and x.is_cuda), "transition tensors must be contiguous float32 CUDA tensors"
        assert tr.env_bins.numel() == self.num_envs and dones.numel() == self.num_envs
        arr_in = (C.c_void_p * 10)(*[x.data_ptr() if x is not None else None for x in ins])
        arr_out = (C.c_void_p * 10)(*[x.data_ptr() for x in outs])
        capi.check(capi.lib().go1_store_transition(arr_in, capi.ptr(dones), capi.ptr(touts), arr_out, capi.ptr(self.dones[t]), self.num_envs,
                                                   self.observations.shape[-1], self.privileged_observations.shape[-1], self.observation_histories.shape[-1],
                                                   self.actions.shape[-1], float(gamma), capi.stream_ptr()), "go1_store_transition")
        self.step += 1

    def clear(self):
        self.step = 0

    def compute_returns(self, last_values, gamma, lam):
        T, N = self.num_transitions_per_env, self.num_envs
        L, st = capi.lib(), capi.stream_ptr()
        last_values = last_values.contiguous()
        capi.check(L.go1_ppo_gae(capi.ptr(self.rewards), capi.ptr(self.dones), capi.ptr(self.values), capi.ptr(last_values),
                                 capi.ptr(self.returns), capi.ptr(self.advantages), capi.ptr(self._stats), T, N, float(gamma), float(lam), st), "gae")
        count = T * N
        if self.process_group is not None:
            import torch.distributed as dist
            dist.all_reduce(self._stats, group=self.process_group)
            count = T * N * dist.get_world_size(self.process_group)
        capi.check(L.go1_ppo_normalize_advantages(capi.ptr(self.advantages), capi.ptr(self._stats), count, T * N, st), "normalize")

    def get_statistics(self):
        done = self.dones
        done[-1] = 1
        flat_dones = done.permute(1, 0, 2).reshape(-1, 1)
        done_indices = torch.cat((flat_dones.new_tensor([-1], dtype=torch.int64), flat_dones.nonzero(as_tuple=False)[:, 0]))
        trajectory_lengths = (done_indices[1:] - done_indices[:-1])
        return trajectory_lengths.float().mean(), self.rewards.mean()

    def gather(self, src, idx, out=None, ldd=None, key=None):
        """out[i] = src.flatten(0,1)[idx[i]] through go1_gather_rows (destination buffers are reused across updates)."""
        flat = src.flatten(0, 1)
        assert idx.dtype == torch.int64 and idx.is_contiguous() and flat.dtype == torch.float32
        w = flat.shape[1]
        ldd = ldd or w
        if out is None and key is not None:
            cache = self.__dict__.setdefault("_gather_bufs", {})
            out = cache.get(key)
            if out is None or out.shape != (idx.shape[0], ldd):
                out = cache[key] = torch.empty(idx.shape[0], ldd, device=flat.device)
        if out is None:
            out = torch.empty(idx.shape[0], ldd, device=flat.device)
        capi.check(capi.lib().go1_gather_rows(capi.ptr(flat), capi.ptr(idx), capi.ptr(out), idx.shape[0], w, ldd, capi.stream_ptr()), "gather")
        return out if ldd == w else out[:, :w]

    def mini_batch_generator(self, num_mini_batches, num_epochs=8, indices=None):
        batch_size = self.num_envs * self.num_transitions_per_env
        mini_batch_size = batch_size // num_mini_batches
        if indices is None:
            indices = torch.randperm(num_mini_batches * mini_batch_size, requires_grad=False, device=self.device)
        dones8 = None
        for epoch in range(num_epochs):
            for i in range(num_mini_batches):
                idx = indices[i * mini_batch_size:(i + 1) * mini_batch_size].contiguous()
                obs = self.gather(self.observations, idx)
                priv_b = self.gather(self.privileged_observations, idx)
                hist_b = self.gather(self.observation_histories, idx, key=("hist", i), ldd=self.hist_pitch)
                w, P = hist_b.shape[1], priv_b.shape[1]
                hist_b.aug = False
                if hist_b.stride(0) >= w + 1 + 2 * P and P >= 1:
                    # spare columns behind the history: [1 | privileged obs | (latent slot)], the augmented inputs of ActorCritic.backward_ppo's
                    # fused first-layer wgrad (bias and trailing-input weight gradients come out of the tensor core with the weight gradient)
                    pad = hist_b.as_strided((hist_b.shape[0], hist_b.stride(0) - w), (hist_b.stride(0), 1), hist_b.storage_offset() + w)
                    pad.zero_()
                    pad[:, 0] = 1.0
                    pad[:, 1:1 + P] = priv_b
                    hist_b.aug = True
                yield (obs, obs, priv_b, hist_b,
                       self.gather(self.actions, idx), self.gather(self.values, idx), self.gather(self.advantages, idx),
                       self.gather(self.returns, idx), self.gather(self.actions_log_prob, idx), self.gather(self.mu, idx),
                       self.gather(self.sigma, idx), dones8, self.gather(self.env_bins, idx))
