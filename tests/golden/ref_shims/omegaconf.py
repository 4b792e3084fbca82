class OmegaConf:
    @staticmethod
    def load(path):
        raise RuntimeError("omegaconf stub")
