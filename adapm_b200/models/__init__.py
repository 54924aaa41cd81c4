"""Model families of the reference's applications, re-designed as batched GPU trainers."""
