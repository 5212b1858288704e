from .is_pos_def import is_pos_def
