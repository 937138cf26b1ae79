class HyperspectralDataset:  # ref umhsnerf/data/utils/hs_dataloader.py:24
    pass
