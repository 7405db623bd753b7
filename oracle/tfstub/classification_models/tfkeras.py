"""`from classification_models.tfkeras import Classifiers` placeholder.  TEST INFRASTRUCTURE ONLY."""


class Classifiers:
    @staticmethod
    def get(name):
        raise NotImplementedError("backbones are out of scope for the stub")
