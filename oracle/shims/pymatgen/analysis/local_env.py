class CrystalNN:
    pass
