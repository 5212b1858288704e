"""Mirror of eeyore/tuners/tuner.py."""


class Tuner:
    def tune(self):
        raise NotImplementedError
