from dataclasses import dataclass


@dataclass
class UMHSDataParserConfig:  # ref umhsnerf/data/umhs_dataparser.py:68-106
    num_classes: int = 5
