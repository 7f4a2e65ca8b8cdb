"""Flag singleton with the reference's names and defaults (config.py:29-123): `opt.<field>`, `opt._parse(kwargs)`.

Only the fields the env path and the ported training loops read are kept; `device` resolves lazily so that importing
the package on a CPU-only box works."""
import warnings


class DefaultConfig(object):
    env = 'RLReachEnv'            # config.py:31
    algo = 'DADDPG_MLP'           # config.py:33
    vis_name = 'Reach_DADDPG'
    vis_port = 8097
    reach_ctr = 0.02              # config.py:41  EE metres per unit action
    reach_dis = 0.01              # config.py:42  reach success distance
    use_gpu = True
    random_seed = 0
    num_episodes = 500
    n_train = 40
    minimal_episodes = 5
    max_steps_one_episode = 500   # config.py:51
    actor_lr = 1e-3
    critic_lr = 1e-3
    hidden_dim = 256
    batch_size = 256
    sigma = 0.1
    tau = 0.005
    gamma = 0.98
    buffer_size = 1000000
    epsilon = 0.01
    target_update = 10
    policy_noise = 0.2
    noise_clip = 0.5
    policy_freq = 3
    q_weight = 0.2
    regularization_weight = 0.005
    her_ratio = 0.8
    # additions of the batched engine
    n_envs = 1
    robot = 'kuka_iiwa'

    @property
    def device(self):
        import torch as t
        return t.device('cuda') if (self.use_gpu and t.cuda.is_available()) else t.device('cpu')

    def _parse(self, kwargs):
        """config.py:81-101: setattr each override, warn on unknown keys, print the user config."""
        for k, v in kwargs.items():
            if not hasattr(self, k):
                warnings.warn("Warning: opt has not attribut %s" % k)
            setattr(self, k, v)
        print('user config:')
        for k in dir(self):
            if not k.startswith('_') and k != 'device':
                print(k, getattr(self, k))


opt = DefaultConfig()
