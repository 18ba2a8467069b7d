"""`python main.py --dataset ml1m --epoch 50 --group 5 --learn sisa --delper 2 --deltype rand`
-- the reference's command line (README.md:33-41), served by ultrare_b200."""
from ultrare_b200.main import main

if __name__ == '__main__':
    main()
