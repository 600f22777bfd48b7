"""C oracle of the Game_of_Flies hot path: test infrastructure only (see gof_oracle.c)."""
