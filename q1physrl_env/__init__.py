"""Drop-in alias of the reference package name: `import q1physrl_env.env` / `q1physrl_env.phys`
give the B200 implementations in `q1physrl_b200` (reference: q1physrl_env/q1physrl_env/)."""
