"""Mirror of eeyore/samplers/sampler.py:1-8."""


class Sampler:
    def draw(self, x, y, savestate=False):
        raise NotImplementedError

    def run(self, num_epochs, num_burnin_epochs, verbose=False, verbose_step=100):
        raise NotImplementedError
